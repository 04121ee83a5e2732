// micro-benchmark: tcgen05.ld throughput per SM (fp32 columns, and 16-bit columns packed two per register)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int PACK>
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32])
{
    if constexpr (PACK == 0)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
    else
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
}

template <int PACK>
__global__ void __launch_bounds__(512, 1) bw_kernel(int iters, long long *cycles, uint32_t *sink)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        uint32_t v[32];
        // 4 loads in flight per warp, columns spread over the 512
        const uint32_t col = (uint32_t)(((warp >> 2) * 128 + (i & 1) * 64) & 511);
        ld32<PACK>(base + col, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 32; e += 8) acc ^= v[e];
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

__global__ void probe_kernel(uint32_t *out)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    if (warp == 0) {
        for (int half = 0; half < 2; half++) {
            uint32_t w[32];
            for (int e = 0; e < 32; e++) w[e] = 0xA000u + (uint32_t)(half * 32 + e);   // column id in the low 16 bits
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
                ::"r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
                  "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]),
                  "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]),
                  "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31]), "r"(base + half * 32)
                : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t v[32];
        ld32<1>(base, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (threadIdx.x == 0) for (int e = 0; e < 32; e++) out[e] = v[e];
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int PACK>
static void run(int warps, const char *name)
{
    long long *d_c, h_c;
    uint32_t *d_s;
    cudaMalloc(&d_c, 8 * 148);
    cudaMalloc(&d_s, 4);
    const int iters = 20000;
    bw_kernel<PACK><<<148, warps * 32>>>(iters, d_c, d_s);
    bw_kernel<PACK><<<148, warps * 32>>>(iters, d_c, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h_c, d_c, 8, cudaMemcpyDeviceToHost);
    const double cols = 32.0 * iters * warps;   // 32 lanes x columns per warp-load
    printf("%-28s warps %2d: %s  %.1f cycles/load/warp  %.1f columns*lanes/clk/SM  %.1f register bytes/clk/SM\n", name, warps,
           cudaGetErrorString(e), (double)h_c / iters, cols * 32 / h_c, 32.0 * 32 * 4 * iters * warps / h_c);
    cudaFree(d_c);
    cudaFree(d_s);
}

int main()
{
    {
        uint32_t *d, h[32];
        cudaMalloc(&d, 128);
        probe_kernel<<<1, 128>>>(d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
        printf("pack::16b probe (%s): v[0]=%08x v[1]=%08x v[15]=%08x v[16]=%08x v[31]=%08x\n", cudaGetErrorString(e), h[0], h[1], h[15], h[16], h[31]);
    }
    for (int w : {4, 8, 16}) run<0>(w, "32x32b.x32 (fp32 columns)");
    for (int w : {4, 8, 16}) run<1>(w, "32x32b.x32.pack::16b");
    return 0;
}
