// Micro-benchmark: FP64 FMA issue rate of ONE warp per SM sub-partition (and of 2, 4) as a function of how many FRESH 64-bit
// register operands an instruction reads.  fp64_issue.cu measures x = fma(x, a, b) with a, b shared by every chain: one fresh
// operand pair per instruction (a and b sit in the operand reuse cache).  The blind rotation's butterflies read two or three
// fresh pairs per DFMA (a constant, two data values).  This kernel keeps 16 independent chains and varies the operand pattern:
//   mode 1: x[u] = fma(x[u], a, b)                 1 fresh pair
//   mode 2: x[u] = fma(x[u], y[u], b)              2 fresh pairs
//   mode 3: x[u] = fma(y[u], z[u], x[u])           3 fresh pairs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operands fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* sink, int iters, double a, double b) {
    constexpr int ILP = 16;
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) { x[u] = threadIdx.x + u; y[u] = 0.999999 + 1e-9 * (threadIdx.x + u); z[u] = 1.0 - 1e-9 * (u + 1); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int u = 0; u < ILP; ++u) {
                if (MODE == 1) x[u] = fma(x[u], a, b);
                else if (MODE == 2) x[u] = fma(x[u], y[u], b);
                else x[u] = fma(y[u], z[u], x[u]);
            }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < ILP; ++u) s += x[u] + y[u] + z[u];
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
    const int sms = p.multiProcessorCount, iters = 4096;
    for (int warps_per_sm : {4, 8, 16}) {
        const int threads = warps_per_sm * 32;
#define RUN(MODE) { float ms = timeit([&] { k<MODE><<<sms, threads>>>(sink, iters, 0.999999, 1e-9); }); \
        double fmas = (double)sms * threads * (double)iters * 16 * 16; \
        printf("warps per sub-partition %d, %d fresh operand pair(s): %.2f TFLOP/s, %.2f cycles per warp-DFMA per sub-partition at 1.965 GHz\n", \
               warps_per_sm / 4, MODE, 2.0 * fmas / ms * 1e-9, ms * 1e-3 * 1.965e9 / ((double)iters * 16 * 16 * (warps_per_sm / 4))); }
        RUN(1) RUN(2) RUN(3)
    }
    return 0;
}
