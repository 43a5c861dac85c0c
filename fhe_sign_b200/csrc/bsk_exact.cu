// bsk_exact.cu — bootstrapping key -> Fourier domain, CORRECTLY ROUNDED (one-off at key upload).
//
// Why.  With l_pbs = 1 and beta = 2^23 the blind rotation's output noise has a large floating-point share: every rounding
// error of the Fourier-domain product lands in the accumulator's MASK polynomial and is amplified by the secret key
// (variance x (1 + N/2) = x 1025 in the phase).  Measured on the CPU (tools/noise_exact_vs_fft.py): a blind rotation with
// exact integer products has 0.81x the variance of the oracle's FFT-based one, and the kernels' FFT-based one had 1.07-1.11x
// when the KEY was transformed with the kernels' own f64 FFT.  A third of that floating-point share is the key's: its
// spectrum carries the forward FFT's rounding error into every product, for every ciphertext, for the key's whole life.
// That part costs nothing to remove: the key is transformed once, so it can be transformed exactly.  With the spectrum
// below the kernels' output variance equals the oracle's (ratio 1.00 on the CPU emulation of the kernels' arithmetic;
// the SURVEY.md 8d gate, tests/test_gpu_pbs.py::test_noise_gate_against_the_oracle_at_full_parameters).
//
// How.  X_k = sum_j (a_j + i a_{j+1024}) zeta^(j (4k+1)), zeta = exp(2 pi i / 4096), k = 0..1023, evaluated directly (no
// FFT) in double-double arithmetic: the 64-bit integer coefficients are split exactly into (hi, lo) doubles, the twiddles
// come from a host table computed in 80-bit long double (64-bit mantissa) and split the same way, products are formed with
// FMA-exact two-products and accumulated with two-sums.  The result is rounded to double once, at the end: relative error
// about 2^-63 before that rounding instead of about 2^-50 after ten butterfly levels.  1 M double-double complex
// multiply-adds per polynomial, 4 n polynomials: tens of milliseconds on a B200, once per key.
// The two output orders are those of launch_bsk_convert (ring kernel: register slot r holds k1 = brev5(r)) and
// launch_bsk_convert_stream (position p of the consumption order holds k1 = freq_at(p)); lane = k2, k = k2 + 32 k1.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "fsc_internal.h"
#include "pbs_core2.cuh"

namespace fsc {

namespace {

struct dd { double h, l; };

__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
// acc += x (x a plain double), acc kept as an unevaluated sum h + l with |l| small
__device__ __forceinline__ void dd_acc(dd& acc, double x) {
    double s, e;
    two_sum(acc.h, x, s, e);
    acc.h = s;
    acc.l += e;
}
// acc += a * b for double-doubles a, b (error terms of second order dropped: 2^-105 relative)
__device__ __forceinline__ void dd_fma(dd& acc, const dd a, const dd b) {
    const double p = a.h * b.h;
    const double e = fma(a.h, b.h, -p);
    dd_acc(acc, p);
    acc.l += e + fma(a.h, b.l, a.l * b.h);
}
__device__ __forceinline__ void dd_norm(dd& a) {
    const double s = a.h + a.l;
    a.l = a.l - (s - a.h);
    a.h = s;
}

// grid = (n * 4 polynomials, 8), block = 128: thread computes frequency k = blockIdx.y * 128 + threadIdx.x
// tw: 4096 x (cos hi, cos lo, sin hi, sin lo)
__global__ void __launch_bounds__(128) bsk_exact_kernel(const uint64_t* __restrict__ bsk, const double4* __restrict__ tw,
                                                        cplx* __restrict__ out_ring, cplx* __restrict__ out_stream) {
    __shared__ dd za[1024], zb[1024];
    const uint64_t* src = bsk + (size_t)blockIdx.x * kN;
    for (int j = threadIdx.x; j < 1024; j += 128) {
        const int64_t a = (int64_t)src[j], b = (int64_t)src[j + 1024];
        dd x, y;
        x.h = (double)a; x.l = (double)(a - (int64_t)__double2ll_rn(x.h));      // |a| <= 2^63: the conversion back is exact except at 2^63 itself
        y.h = (double)b; y.l = (double)(b - (int64_t)__double2ll_rn(y.h));
        if (x.h >= 9223372036854775808.0) x.l = (double)(a - INT64_MAX) - 1.0;   // a rounded up to 2^63: remainder a - 2^63
        if (y.h >= 9223372036854775808.0) y.l = (double)(b - INT64_MAX) - 1.0;
        za[j] = x; zb[j] = y;
    }
    __syncthreads();
    const int k = blockIdx.y * 128 + threadIdx.x;
    const unsigned step = (unsigned)(4 * k + 1);
    dd re = {0.0, 0.0}, im = {0.0, 0.0};
    unsigned e = 0;
    for (int j = 0; j < 1024; ++j, e = (e + step) & 4095u) {
        const double2 tc = __ldg(reinterpret_cast<const double2*>(tw) + 2 * e), ts = __ldg(reinterpret_cast<const double2*>(tw) + 2 * e + 1);
        const dd c = {tc.x, tc.y}, s = {ts.x, ts.y}, ms = {-ts.x, -ts.y};
        const dd a = za[j], b = zb[j];
        dd_fma(re, a, c); dd_fma(re, b, ms);      // (a + i b)(c + i s) = (a c - b s) + i (a s + b c)
        dd_fma(im, a, s); dd_fma(im, b, c);
        if ((j & 15) == 15) { dd_norm(re); dd_norm(im); }
    }
    dd_norm(re); dd_norm(im);
    cplx v; v.x = re.h; v.y = im.h;
    const int k2 = k & 31, k1 = k >> 5;
    const size_t i = blockIdx.x >> 2, g = blockIdx.x & 3;
    if (out_ring) out_ring[((i * 32 + brev5(k1)) * 4 + g) * 32 + k2] = v;
    if (out_stream) out_stream[((i * 32 + freq_pos(k1)) * 4 + g) * 32 + k2] = v;
}

const double4* exact_twiddles() {      // device pointer, built once per device
    static double4* per_dev[64] = {};
    static std::mutex mu;
    int dev = 0;
    FSC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    double4*& d = per_dev[dev & 63];
    if (!d) {
        std::vector<double4> h(4096);
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int e = 0; e < 4096; ++e) {
            long double c = cosl(two_pi * (long double)e / 4096.0L), s = sinl(two_pi * (long double)e / 4096.0L);
            if ((e & 1023) == 0) {      // exact axis values
                const int q = e >> 10;
                c = (q == 0) ? 1.0L : (q == 2) ? -1.0L : 0.0L;
                s = (q == 1) ? 1.0L : (q == 3) ? -1.0L : 0.0L;
            }
            h[e].x = (double)c; h[e].y = (double)(c - (long double)h[e].x);
            h[e].z = (double)s; h[e].w = (double)(s - (long double)h[e].z);
        }
        FSC_CUDA_CHECK(cudaMalloc(&d, 4096 * sizeof(double4)));
        FSC_CUDA_CHECK(cudaMemcpy(d, h.data(), 4096 * sizeof(double4), cudaMemcpyHostToDevice));
    }
    return d;
}

}  // namespace

// out_ring / out_stream: either may be null
void launch_bsk_convert_exact(const uint64_t* bsk_std, void* out_ring, void* out_stream, int n, cudaStream_t st) {
    const double4* tw = exact_twiddles();
    bsk_exact_kernel<<<dim3((unsigned)(n * 4), 8), 128, 0, st>>>(bsk_std, tw, reinterpret_cast<cplx*>(out_ring),
                                                                 reinterpret_cast<cplx*>(out_stream));
}

}  // namespace fsc
