// Micro-benchmark: does the FP64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA) run on its own pipe on B200, or does it share the
// DFMA units?  north_star: "tensor cores are used only if ... ncu shows it beating the FP64 ... CUDA-core path".  tcgen05 has no
// FP64 operand kind, so DMMA is the only tensor instruction the blind rotation could use.  Four kernels, same launch shape:
//   mode 0: 16 independent DFMA chains per thread                     (64 flop per warp instruction)
//   mode 1: 8 independent DMMA accumulator chains per warp            (512 flop per warp instruction)
//   mode 2: both in every warp, one DMMA per 8 DFMAs (equal flops)    -> adds up only if the pipes are separate
//   mode 3: even warps DFMA only, odd warps DMMA only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_dmma fp64_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void k(double* sink, int iters, double a, double b) {
    double x[16], c[8][2];
#pragma unroll
    for (int u = 0; u < 16; ++u) x[u] = threadIdx.x + u;
#pragma unroll
    for (int u = 0; u < 8; ++u) { c[u][0] = threadIdx.x; c[u][1] = u; }
    const bool odd = (threadIdx.x >> 5) & 1;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0 || MODE == 2 || (MODE == 3 && !odd)) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int u = 0; u < 16; ++u) x[u] = fma(x[u], a, b);
        }
        if (MODE == 1 || MODE == 2 || (MODE == 3 && odd)) {
#pragma unroll
            for (int u = 0; u < 8; ++u) dmma(c[u], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 16; ++u) s += x[u];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
    if (s == 12345.678) sink[0] = s;
}
template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* sink; cudaMalloc(&sink, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 8192;
    printf("# tools/ubench/fp64_dmma (%s, %d SMs): DFMA = 64 flop / warp instruction, DMMA m8n8k4 = 512\n", p.name, sms);
    for (int warps_per_sm : {4, 8, 16}) {
        const int threads = warps_per_sm * 32;
        const double warps = (double)sms * warps_per_sm;
#define RUN(MODE, DFMA_WARPS, DMMA_WARPS) { float ms = timeit([&] { k<MODE><<<sms, threads>>>(sink, iters, 0.999999, 1e-9); }); \
        const double f_fma = warps * (DFMA_WARPS) * iters * 64.0 * 64.0, f_mma = warps * (DMMA_WARPS) * iters * 8.0 * 512.0; \
        printf("warps/SM %2d mode %d: %.3f ms  DFMA %.2f TFLOP/s + DMMA %.2f TFLOP/s = %.2f TFLOP/s\n", warps_per_sm, MODE, ms, \
               f_fma / ms * 1e-9, f_mma / ms * 1e-9, (f_fma + f_mma) / ms * 1e-9); }
        RUN(0, 1.0, 0.0) RUN(1, 0.0, 1.0) RUN(2, 1.0, 1.0) RUN(3, 0.5, 0.5)
    }
    return 0;
}
