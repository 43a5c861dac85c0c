// Micro-benchmark: does ONE instruction stream that holds two independent pieces of work — FP64 butterflies on 16 points
// (split_levels25, 192 FP64 instructions) and a 16-value shared-memory transpose (16 STS.128 + 16 LDS.128) or an integer block —
// take max(A, B) or A + B cycles?  8 warps per SM (two per sub-partition), every warp the same code: the lock-step situation of
// the blind-rotation kernels.  Premise test for a kernel whose warps carry two half-polynomials each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../fhe_sign_b200/csrc -o interleave interleave.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "pbs_core3.cuh"
using namespace fsc;

template <int MODE>      // 1: butterflies only, 2: transpose only, 3: both in one basic block, 4: integer block only, 5: butterflies + integer block
__global__ void __launch_bounds__(256, 1) k(double* sink, long long* cyc, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    cplx* tab = reinterpret_cast<cplx*>(smem);                       // 16 x 32 per-lane constants
    cplx* T = tab + 512 + (threadIdx.x >> 5) * (32 * 17);            // per warp: [32][17] complex
    for (int t = threadIdx.x; t < 512; t += blockDim.x) { tab[t].x = 0.7 + 1e-3 * t; tab[t].y = 0.3 - 1e-4 * t; }
    __syncthreads();
    const int lane = threadIdx.x & 31, h = (threadIdx.x >> 7) & 1;
    cplx a[16], b[16];
    int q[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { a[j].x = 1e-3 * (threadIdx.x + j); a[j].y = 1.0 - 1e-3 * j; b[j] = a[j]; q[j] = threadIdx.x * 7 + j; }
    const StridedConsts sp{tab + lane, 32};
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        if (MODE == 1 || MODE == 3 || MODE == 5) {
            split_levels25(h, sp, a);
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j].x *= 1e-3;
        }
        if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int j = 0; j < 16; ++j) T[j * 17 + (lane & 15)] = b[j];       // row store (half a warp per row: 2-way, like a 16-wide transpose)
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; ++j) b[j] = T[(lane & 15) * 17 + j];
            __syncwarp();
        }
        if (MODE == 4 || MODE == 5) {
#pragma unroll
            for (int r = 0; r < 12; ++r)
#pragma unroll
                for (int j = 0; j < 16; ++j) q[j] = (q[j] * 3 + (q[(j + 1) & 15] >> 2)) ^ r;      // 16 x 12 x ~3 integer instructions
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j].x + a[j].y + b[j].x + b[j].y + q[j];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* sink; long long* cyc; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 2000;
    const size_t smem = 512 * 16 + 8 * 32 * 17 * 16;
    printf("# tools/ubench/interleave (%s): cycles per loop trip, 8 warps per SM, every warp the same code\n", p.name);
    long long h;
#define RUN(M, what) { cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<M><<<p.multiProcessorCount, 256, smem>>>(sink, cyc, iters); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-58s %6.0f cycles\n", what, (double)h / iters); }
    RUN(1, "A: butterflies (192 + 16 FP64 instructions)")
    RUN(2, "B: transpose of 16 complex values through shared memory")
    RUN(3, "A + B in one basic block")
    RUN(4, "C: integer block (about 580 ALU / FMA-pipe instructions)")
    RUN(5, "A + C in one basic block")
    return 0;
}
