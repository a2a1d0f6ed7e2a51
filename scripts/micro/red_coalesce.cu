// Microbenchmark: how does the cost of fp64 global reductions depend on how a warp issues them?
//   A: 5 lanes x 6 RED, component-major histogram [6][nb]      (pb2_xi_cross_chunk before)
//   B: 5 lanes x 6 RED, bin-major histogram [nb][8]             (pb2_xi_auto_diag)
//   C: 30 lanes x 1 RED, bin-major: 6 adjacent lanes share one 64-byte line
//   D: 30 lanes x 1 RED, component-major (30 distinct lines)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_coalesce red_coalesce.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red(double *p, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned hash(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <int MODE>
__global__ void k(double *h, int nb, int iters, int groups)
{
    const int lane = threadIdx.x & 31;
    const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 1) {
            if (lane % 7 == 0 && lane / 7 < groups) {
                const int bin = hash(w * 7919u + it * 31u + lane) % nb;
#pragma unroll
                for (int m = 0; m < 6; m++)
                    red(MODE == 0 ? h + (size_t)m * nb + bin : h + (size_t)bin * 8 + m, 1.0);
            }
        } else {
            const int g = lane / 6, m = lane % 6;
            if (g < groups) {
                const int bin = hash(w * 7919u + it * 31u + g * 7) % nb;
                red(MODE == 3 ? h + (size_t)m * nb + bin : h + (size_t)bin * 8 + m, 1.0);
            }
        }
    }
}
int main()
{
    const int nb = 5000 * 64;  // 64 HEALPix rows of 5000 bins in flight: 20 MB, L2-resident
    double *h;
    cudaMalloc(&h, (size_t)nb * 8 * sizeof(double));
    cudaMemset(h, 0, (size_t)nb * 8 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4000, grid = 148 * 4, block = 256;
    for (int groups = 1; groups <= 5; groups += 2)
        for (int mode = 0; mode < 4; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<grid, block>>>(h, nb, iters, groups);
                if (mode == 1) k<1><<<grid, block>>>(h, nb, iters, groups);
                if (mode == 2) k<2><<<grid, block>>>(h, nb, iters, groups);
                if (mode == 3) k<3><<<grid, block>>>(h, nb, iters, groups);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep == 1) {
                    const double reds = (double)grid * (block / 32) * iters * groups * 6;
                    printf("groups %d mode %c: %.3f ms, %.3e red/s, %.3e runs/s\n", groups, "ABCD"[mode],
                           ms, reds / (ms * 1e-3), reds / 6 / (ms * 1e-3));
                }
            }
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
