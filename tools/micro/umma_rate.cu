// micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16, fp32 accumulate), both operands from shared memory (SS),
// M = 128, N = 128 or 256, one CTA per SM -- is the 128 x 128 SS tile shared-memory-bandwidth bound?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../gorse_b200/csrc -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "umma.cuh"
using namespace gb::mma;

// MODE 0: MMAs alone; 1: 8 warps stream LDS.128; 2: 16 warps stream tcgen05.ld from the idle accumulator half;
// 3: 16 warps poll an mbarrier that never flips; 4: 16 warps run an FMNMX3 stream (issue-slot pressure only)
template <int N, int MODE>
__global__ void __launch_bounds__(128 + 32 * 16, 1) rate_kernel(int iters, long long *cycles, float *sink)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *a = smem, *b = smem + 2 * 16384;                   // A: 2 k-blocks [128 x 64] bf16; B: 2 k-blocks [N x 64]
    float *junk = reinterpret_cast<float *>(b + 2 * N * 128);    // 32 KB the loader warps read
    uint64_t *bar = reinterpret_cast<uint64_t *>(junk + 8192);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 2);
    volatile uint32_t *stop = slot + 1;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (2 * 16384 + 2 * N * 128 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1) {
        const uint64_t da = umma_desc(a, 0), db = umma_desc(b, 0);
        const uint32_t a_hi = (uint32_t)(da >> 32), b_hi = (uint32_t)(db >> 32), a_lo = (uint32_t)da, b_lo = (uint32_t)db;
        auto desc = [](uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; };
        const long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; i++) {
#pragma unroll
                for (uint32_t kb = 0; kb < 2; kb++)
#pragma unroll
                    for (uint32_t k = 0; k < 4; k++)
                        umma_bf16(tmem, desc(a_lo + kb * 1024 + k * 2, a_hi), desc(b_lo + kb * (N * 8) + k * 2, b_hi), IDESC, (kb | k) != 0);
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) { cycles[blockIdx.x] = t1 - t0; *stop = 1; mbar_arrive(bar + 1); }
    } else if (warp >= 4 && MODE != 0) {
        const int ew = warp - 4;
        float acc = 0.f;
        if (MODE == 1) {
            if (ew < 8) {
                const float4 *p = reinterpret_cast<const float4 *>(junk);
                for (int i = 0; !*stop; i++) {
                    const float4 v = p[(i * 32 + (threadIdx.x & 31)) & 2047];
                    acc += v.x + v.y + v.z + v.w;
                }
            }
        } else if (MODE == 2) {
            const uint32_t base = tmem + ((uint32_t)((ew & 3) * 32) << 16) + 256u + (uint32_t)((ew >> 2) * 64);
            for (int i = 0; !*stop; i++) {
                uint32_t v[32];
                tmem_ld32(base + (uint32_t)((i & 1) * 32), v);
                acc += __uint_as_float(v[0] ^ v[31]);
            }
        } else if (MODE == 3) {
            mbar_wait(bar + 1, 0);
        } else {
            float x = (float)threadIdx.x, y = 1.f, z = 2.f;
            for (int i = 0; !*stop; i++) {
#pragma unroll
                for (int r = 0; r < 16; r++) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z)); }
            }
            acc = x;
        }
        if (acc == 12345.f) sink[0] = acc;
    }
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int MODE>
static void run()
{
    static const char *names[] = {"MMAs alone", "8 warps stream LDS.128", "16 warps stream tcgen05.ld", "16 warps poll an mbarrier", "16 warps run FMNMX3"};
    long long *d_c, h_c[148];
    float *d_s;
    cudaMalloc(&d_c, 8 * 148);
    cudaMalloc(&d_s, 4);
    const int iters = 20000;
    const size_t sm = 1024 + 2 * 16384 + 2 * N * 128 + 32768 + 64;
    cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    rate_kernel<N, MODE><<<148, 128 + 32 * 16, sm>>>(iters, d_c, d_s);
    rate_kernel<N, MODE><<<148, 128 + 32 * 16, sm>>>(iters, d_c, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h_c, d_c, 8 * 148, cudaMemcpyDeviceToHost);
    const double cyc = (double)h_c[0] / (iters * 8.0);
    printf("M128 N%3d K16 SS, %-28s: %s  %.1f cycles per MMA (floor %d)  = %.0f%% of the tensor peak\n", N, names[MODE], cudaGetErrorString(e), cyc,
           N / 2, 100.0 * (N / 2) / cyc);
    cudaFree(d_c);
    cudaFree(d_s);
}

int main()
{
    run<128, 0>();
    run<256, 0>();
    run<128, 1>();
    run<128, 2>();
    run<128, 3>();
    run<128, 4>();
    run<256, 1>();
    run<256, 2>();
    return 0;
}
