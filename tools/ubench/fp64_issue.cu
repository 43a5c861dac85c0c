// Micro-benchmark: FP64 FMA issue rate per SM as a function of resident warps per scheduler and ILP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* sink, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) x[u] = threadIdx.x + u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int u = 0; u < ILP; ++u) x[u] = fma(x[u], a, b);
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < ILP; ++u) s += x[u];
    if (s == 12345.678) sink[0] = s;
}
// mixed: FP64 chains + INT work in the same warp
template <int ILP>
__global__ void kmix(double* sink, int iters, double a, double b) {
    double x[ILP]; unsigned y[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) { x[u] = threadIdx.x + u; y[u] = threadIdx.x * 7 + u; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int u = 0; u < ILP; ++u) { x[u] = fma(x[u], a, b); y[u] = (y[u] ^ (y[u] >> 3)) + 0x9e3779b9u; }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int u = 0; u < ILP; ++u) { s += x[u]; t += y[u]; }
    if (s == 12345.678 || t == 0x12345) sink[0] = s;
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
    for (int warps_per_sm : {4, 8, 16, 32}) {
        const int threads = warps_per_sm * 32;
#define RUN(ILP) { float ms = timeit([&] { k<ILP><<<sms, threads>>>(sink, iters, 0.999999, 1e-9); }); \
        double fl = 2.0 * sms * threads * (double)iters * 16 * ILP; \
        float ms2 = timeit([&] { kmix<ILP><<<sms, threads>>>(sink, iters, 0.999999, 1e-9); }); \
        printf("warps/SM %2d ILP %2d: %.2f TFLOP/s   (with 2 INT ops per FMA: %.2f TFLOP/s)\n", warps_per_sm, ILP, fl / ms * 1e-9, fl / ms2 * 1e-9); }
        RUN(1) RUN(2) RUN(4) RUN(8) RUN(16)
    }
    return 0;
}
