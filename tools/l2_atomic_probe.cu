// l2_atomic_probe.cu -- microbenchmark behind DESIGN.md "hot rows": how fast can one B200 L2 slice absorb
// red.global.add.v4.f32 (and ld.cg) to a single 256-byte row, and does striping the row's 16-byte pieces
// over different 256-byte blocks (different L2 slices) help?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_atomic_probe l2_atomic_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red4(float *p, float4 v)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// mode 0: red.v4 to piece (tid%16) of one row; piece stride in floats = stride
// mode 1: ld.cg of the same pieces
// mode 2: scalar red.f32 to the 64 floats of the row
// mode 3: red.v4 to pieces of `nrows` different rows chosen round-robin (spread)
__global__ void probe(float *buf, int mode, long stride, int iters, int nrows, long row_stride, float *sink)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int piece = t & 15;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; it++) {
        long row = nrows > 1 ? ((t >> 4) + it) % nrows : 0;
        float *p = buf + row * row_stride + piece * stride;
        if (mode == 0 || mode == 3) red4(p, make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f));
        else if (mode == 1) { float4 v = __ldcg((const float4 *)p); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        else { for (int k = 0; k < 4; k++) atomicAdd(p + k, 1e-9f); }
    }
    if (acc.x == 12345.f) sink[0] = acc.x + acc.y + acc.z + acc.w;
}

int main()
{
    float *buf, *sink;
    size_t bytes = (size_t)1 << 30;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 16);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256, iters = 200;
    const double ops = (double)grid * block * iters;
    struct Case { const char *name; int mode; long stride; int nrows; long row_stride; } cases[] = {
        {"red.v4  one 256B row (pieces contiguous)", 0, 4, 1, 0},
        {"red.v4  one row, pieces striped 256 B apart", 0, 64, 1, 0},
        {"red.v4  one row, pieces striped 4 KB apart", 0, 1024, 1, 0},
        {"red.v4  one row, pieces striped 1 MB apart", 0, 262144, 1, 0},
        {"ld.cg   one 256B row", 1, 4, 1, 0},
        {"ld.cg   one row, pieces striped 4 KB apart", 1, 1024, 1, 0},
        {"red.f32 x4 one 256B row", 2, 4, 1, 0},
        {"red.v4  16 rows 256 B apart (contiguous table)", 3, 4, 16, 64},
        {"red.v4  1024 rows (contiguous table)", 3, 4, 1024, 64},
        {"red.v4  100K rows (25.6 MB table, L2)", 3, 4, 100000, 64},
        {"red.v4  1M rows (256 MB table, HBM)", 3, 4, 1000000, 64},
    };
    for (auto &c : cases) {
        probe<<<grid, block>>>(buf, c.mode, c.stride, 10, c.nrows, c.row_stride, sink);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        probe<<<grid, block>>>(buf, c.mode, c.stride, iters, c.nrows, c.row_stride, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-52s %8.3f ms  %8.2f G 16B-ops/s  (%6.1f GB/s)\n", c.name, ms, ops / ms * 1e-6, ops * 16 / ms * 1e-6);
    }
    return 0;
}
