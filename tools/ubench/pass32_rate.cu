// Micro-benchmark: cycles per 32-point pass (pbs_core2.cuh pass32, 512 FP64 instructions) for 1, 2 warps per SM sub-partition,
// per-lane constants in shared memory (passes 1 / 3 of the blind rotation) against a uniform table (passes 0 / 2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../fhe_sign_b200/csrc -o pass32_rate pass32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "pbs_core2.cuh"
using namespace fsc;

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* sink, long long* cyc, int iters) {
    __shared__ cplx tab[16 * 32];
    for (int t = threadIdx.x; t < 512; t += blockDim.x) { tab[t].x = 0.7 + 1e-3 * t; tab[t].y = 0.3 - 1e-4 * t; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    cplx v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { v[j].x = 1e-3 * (threadIdx.x + j); v[j].y = 1.0 - 1e-3 * j; }
    const StridedConsts sp{MODE == 0 ? tab + lane : tab, MODE == 0 ? 32 : 1};
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        pass32(v, sp);
#pragma unroll
        for (int j = 0; j < 32; ++j) { v[j].x *= 1e-3; }      // keep the values bounded (32 extra DMUL per pass)
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += v[j].x + v[j].y;
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* sink; long long* cyc; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 2000;
    printf("# tools/ubench/pass32_rate (%s): cycles per pass32 (512 FP64 + 32 scaling DMUL), per warp\n", p.name);
    for (int warps : {4, 8}) {
        long long h;
        k<0><<<p.multiProcessorCount, warps * 32>>>(sink, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps per sub-partition %d, per-lane constants: %.0f cycles per pass (%.2f per FP64 instruction of the sub-partition)\n", warps / 4, (double)h / iters, (double)h / iters / (544.0 * (warps / 4)));
        k<1><<<p.multiProcessorCount, warps * 32>>>(sink, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps per sub-partition %d, uniform constants:  %.0f cycles per pass (%.2f per FP64 instruction of the sub-partition)\n", warps / 4, (double)h / iters, (double)h / iters / (544.0 * (warps / 4)));
    }
    return 0;
}
