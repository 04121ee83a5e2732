// micro-benchmark: the accumulator hand-off of topk_mma_kernel without TMA and without epilogue work.
// MMA warp: wait t_empty[as] -> 16 MMAs (2 accumulators x 8) -> commit t_full[as]; E epilogue warps: wait t_full[as] ->
// (optionally tcgen05.ld) -> arrive t_empty[as].  What is the period per step against the 1024-cycle MMA floor?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../gorse_b200/csrc -o umma_pipe umma_pipe.cu && ./umma_pipe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "umma.cuh"
using namespace gb::mma;

// MODE 0: no hand-off at all (commits to barriers nobody reads); 1: hand-off, E warps; 2: hand-off + two tcgen05.ld per warp
template <int MODE, int E, int NSTAGE>
__global__ void __launch_bounds__(128 + 32 * 16, 1) pipe_kernel(int steps, long long *cycles, float *sink)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *a = smem, *b = smem + 4 * 16384;                   // A: 2 tiles x 2 k-blocks; B: 2 k-blocks
    uint64_t *bars = reinterpret_cast<uint64_t *>(b + 2 * 16384);
    uint64_t *t_full = bars, *t_empty = bars + 4, *dummy = bars + 8;
    uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 12);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (6 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; s++) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], E); mbar_init(&dummy[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    constexpr uint32_t N = 256 / NSTAGE * 2 / 2;   // NSTAGE 2: two 128-column accumulators per stage; NSTAGE 4: two 64-column ones
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1) {
        const uint64_t da = umma_desc(a, 0), db = umma_desc(b, 0);
        const uint32_t a_hi = (uint32_t)(da >> 32), b_hi = (uint32_t)(db >> 32), a_lo = (uint32_t)da, b_lo = (uint32_t)db;
        auto desc = [](uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; };
        const long long t0 = clock64();
        for (uint32_t at = 0; at < (uint32_t)steps; at++) {
            const uint32_t as = at % NSTAGE;
            if (MODE != 0) mbar_wait(&t_empty[as], ((at / NSTAGE) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
#pragma unroll
                for (uint32_t m = 0; m < 2; m++)
#pragma unroll
                    for (uint32_t kb = 0; kb < 2; kb++)
#pragma unroll
                        for (uint32_t k = 0; k < 4; k++)
                            umma_bf16(tmem + (m * NSTAGE + as) * N, desc(a_lo + (m * 2 + kb) * 1024 + k * 2, a_hi), desc(b_lo + kb * 1024 + k * 2, b_hi),
                                      IDESC, (kb | k) != 0);
                umma_commit(&dummy[as]);
                umma_commit(MODE != 0 ? &t_full[as] : &dummy[2 + (as & 1)]);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&dummy[0]);
        __syncwarp();
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    } else if (warp >= 4 && warp < 4 + E && MODE != 0) {
        const int ew = warp - 4;
        float acc = 0.f;
        const uint32_t base = tmem + ((uint32_t)((ew & 3) * 32) << 16);
        for (uint32_t at = 0; at < (uint32_t)steps; at++) {
            const uint32_t as = at % NSTAGE;
            mbar_wait(&t_full[as], (at / NSTAGE) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (MODE == 2) {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(base + (uint32_t)(((ew >> 3) * NSTAGE + as) * N + ((ew >> 2) & 1) * (N / 2)), v0);
                tmem_ld32_issue(base + (uint32_t)(((ew >> 3) * NSTAGE + as) * N + ((ew >> 2) & 1) * (N / 2) + (N >= 128 ? 32 : 0)), v1);
                tmem_ld_wait();
                acc += __uint_as_float(v0[0] ^ v1[31]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[as]);
        }
        if (acc == 12345.f) sink[0] = acc;
    }
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE, int E, int NSTAGE>
static void run(const char *name)
{
    long long *d_c, h_c[148];
    float *d_s;
    cudaMalloc(&d_c, 8 * 148);
    cudaMalloc(&d_s, 4);
    const int steps = 20000;
    const size_t sm = 1024 + 6 * 16384 + 256;
    cudaFuncSetAttribute(pipe_kernel<MODE, E, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    pipe_kernel<MODE, E, NSTAGE><<<148, 128 + 32 * 16, sm>>>(steps, d_c, d_s);
    pipe_kernel<MODE, E, NSTAGE><<<148, 128 + 32 * 16, sm>>>(steps, d_c, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h_c, d_c, 8 * 148, cudaMemcpyDeviceToHost);
    const double floor_ = NSTAGE == 2 ? 1024.0 : 512.0;
    printf("%-64s: %s  %.0f cycles per step (MMA floor %.0f) = %.0f%%\n", name, cudaGetErrorString(e), (double)h_c[0] / steps, floor_,
           100.0 * floor_ * steps / h_c[0]);
    cudaFree(d_c);
    cudaFree(d_s);
}

int main()
{
    run<0, 16, 2>("no hand-off, 2 commits per 16 MMAs");
    run<1, 1, 2>("hand-off with 1 warp, 2 stages x (2 x 128 columns)");
    run<1, 16, 2>("hand-off with 16 warps, 2 stages");
    run<2, 16, 2>("hand-off with 16 warps + 2 tcgen05.ld each, 2 stages");
    run<1, 16, 4>("hand-off with 16 warps, 4 stages x (2 x 64 columns)");
    run<2, 16, 4>("hand-off with 16 warps + tcgen05.ld, 4 stages x (2 x 64 columns)");
    return 0;
}
