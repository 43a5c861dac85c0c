// pbs_core.cuh — per-lane building blocks of the warp-resident blind rotation.
//
// One warp owns one ciphertext.  A CMUX step  acc += GGSW_i (x) (X^a acc - acc)  is, per warp:
//   head      : rotated difference + signed gadget decomposition (l = 1) of both polynomials
//   dft32 x2  : negacyclic FFT, N = 2048 reals -> M = 1024 complex, as 32 x 32:
//               pass 1 (lane j1, registers j2) -> swizzled shared-memory transpose ->
//               pass 2 (lane k2, registers j1); every pass evaluates a degree-<32 polynomial at the
//               32 roots of x^32 = zeta^(32 g) by splitting x^2h - c = (x^h - s)(x^h + s)
//   mac       : (k+1) x (k+1) Fourier-domain product with the bootstrapping-key GGSW, in registers
//   inverse   : the exact reverse (Gentleman-Sande butterflies, conjugate constants)
//   tail      : f64 -> torus rounding and accumulation
// tools/fft_proto.py is the numpy statement of the same index conventions.
//
// Everything here is __host__ __device__ and takes `lane` as an argument so that
// tests/emu/pbs_emu.cpp can execute the very same code lane by lane on the CPU.
//
// Replaces (concept): tfhe 0.10.0 core_crypto blind_rotate_assign / add_external_product_assign
// (Cargo.lock:482-485; reached from every FheUint operator in src/biguint.rs:110-248).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define FSC_HD __host__ __device__ __forceinline__
#define FSC_ALIGN16 __align__(16)
#else
#define FSC_HD inline
#define FSC_ALIGN16 alignas(16)
#endif

// Compiler-only fence used inside the long unrolled load/convert loops: it stops ptxas from hoisting every
// shared-memory load of the loop to its top, which would keep all of their destinations live at once.
#if defined(__CUDA_ARCH__)
#define FSC_SCHED_FENCE(i, every) do { if (FSC_FENCE_EVERY > 0 && ((i) % (every)) == (every) - 1) asm volatile("" ::: "memory"); } while (0)
#else
#define FSC_SCHED_FENCE(i, every) do { } while (0)
#endif
#ifndef FSC_FENCE_EVERY
#define FSC_FENCE_EVERY 0
#endif

namespace fsc {

constexpr int kN = 2048;        // polynomial size (fixed by the kernel design)
constexpr int kM = 1024;        // complex points
constexpr int kLogN2 = 12;      // log2(2N)

struct FSC_ALIGN16 cplx { double x, y; };

template <typename T> struct FSC_ALIGN16 pair_t { T x, y; };
template <> struct alignas(8) pair_t<uint32_t> { uint32_t x, y; };

FSC_HD constexpr int brev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// exp(2 pi i e / 4096)
FSC_HD cplx twiddle4096(int e) {
    e &= 4095;
    cplx r;
#if defined(__CUDA_ARCH__)
    sincospi((double)e * (1.0 / 2048.0), &r.y, &r.x);
#else
    // exact octant reduction keeps the host tables as accurate as sincospi
    const long double a = 3.14159265358979323846264338327950288L * (long double)e / 2048.0L;
    r.x = (double)cosl(a); r.y = (double)sinl(a);
    if (e == 0) { r.x = 1; r.y = 0; } else if (e == 1024) { r.x = 0; r.y = 1; }
    else if (e == 2048) { r.x = -1; r.y = 0; } else if (e == 3072) { r.x = 0; r.y = -1; }
#endif
    return r;
}

// The 16 stored node constants of a pass with root parameter g (x^32 = zeta^(32 g)):
// index 0 -> level 1; index 2^(L-2)+t -> level L>=2, even node 2t (odd node 2t+1 = i * even node).
FSC_HD constexpr int node_level(int ci) { return ci >= 8 ? 5 : ci >= 4 ? 4 : ci >= 2 ? 3 : ci == 1 ? 2 : 1; }
FSC_HD constexpr int node_exponent(int ci, int g) {
    return (32 >> node_level(ci)) * g + 64 * brev5(2 * (node_level(ci) == 1 ? 0 : ci - (1 << (node_level(ci) - 2))));
}

FSC_HD void lane_consts(int g, cplx (&s)[16]) {
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) s[ci] = twiddle4096(node_exponent(ci, g));
}
constexpr int kP1G = 32;      // root parameter of pass 1 (uniform): x^32 = zeta^(32 * 32) = i

// ---- 32-point passes ------------------------------------------------------------------
// SP::get(ci) returns stored constant ci.
template <class SP>
FSC_HD void dft32_fwd(cplx (&v)[32], const SP& sp) {
#pragma unroll
    for (int L = 1; L <= 5; ++L) {
        const int half = 16 >> (L - 1);
#pragma unroll
        for (int m = 0; m < (1 << (L - 1)); ++m) {
            const int base = m * 2 * half;
            const int ci = (L == 1) ? 0 : ((1 << (L - 2)) + (m >> 1));
            const bool odd = (L > 1) && (m & 1);
            const cplx s = sp.get(ci);
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const cplx lo = v[base + j], hi = v[base + half + j];
                const double tx = s.x * hi.x - s.y * hi.y;
                const double ty = s.x * hi.y + s.y * hi.x;
                if (!odd) {
                    v[base + j].x = lo.x + tx;        v[base + j].y = lo.y + ty;
                    v[base + half + j].x = lo.x - tx; v[base + half + j].y = lo.y - ty;
                } else {   // constant is i*s
                    v[base + j].x = lo.x - ty;        v[base + j].y = lo.y + tx;
                    v[base + half + j].x = lo.x + ty; v[base + half + j].y = lo.y - tx;
                }
            }
        }
    }
}

// exact reverse of dft32_fwd up to a factor 32
template <class SP>
FSC_HD void dft32_inv(cplx (&v)[32], const SP& sp) {
#pragma unroll
    for (int L = 5; L >= 1; --L) {
        const int half = 16 >> (L - 1);
#pragma unroll
        for (int m = 0; m < (1 << (L - 1)); ++m) {
            const int base = m * 2 * half;
            const int ci = (L == 1) ? 0 : ((1 << (L - 2)) + (m >> 1));
            const bool odd = (L > 1) && (m & 1);
            const cplx s = sp.get(ci);
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const cplx u = v[base + j], w = v[base + half + j];
                const double dx = u.x - w.x, dy = u.y - w.y;
                v[base + j].x = u.x + w.x; v[base + j].y = u.y + w.y;
                const double ex = s.x * dx + s.y * dy;       // conj(s) * d
                const double ey = s.x * dy - s.y * dx;
                if (!odd) { v[base + half + j].x = ex; v[base + half + j].y = ey; }
                else      { v[base + half + j].x = ey; v[base + half + j].y = -ex; }   // -i * conj(s) * d
            }
        }
    }
}

struct RegConsts {
    const cplx (&s)[16];
    FSC_HD explicit RegConsts(const cplx (&a)[16]) : s(a) {}
    FSC_HD cplx get(int ci) const { return s[ci]; }
};
struct PtrConsts {
    const cplx* s;
    FSC_HD cplx get(int ci) const { return s[ci]; }
};

// ---- swizzled 32x32 transpose buffer: element (row k2, col j1) at k2*32 + (j1 ^ k2) ----
FSC_HD int xaddr(int k2, int j1) { return k2 * 32 + (j1 ^ k2); }

// after forward pass 1 (lane = j1, v[pos] = Y[j1][k2 = brev5(pos)])
FSC_HD void xpose_store_fwd(int lane, cplx* xbuf, const cplx (&v)[32]) {
#pragma unroll
    for (int pos = 0; pos < 32; ++pos) xbuf[xaddr(brev5(pos), lane)] = v[pos];
}
// before forward pass 2 (lane = k2, v[j1])
FSC_HD void xpose_load_fwd(int lane, const cplx* xbuf, cplx (&v)[32]) {
#pragma unroll
    for (int j1 = 0; j1 < 32; ++j1) v[j1] = xbuf[xaddr(lane, j1)];
}
// after inverse pass 2 (lane = k2, v[j1])
FSC_HD void xpose_store_inv(int lane, cplx* xbuf, const cplx (&v)[32]) {
#pragma unroll
    for (int j1 = 0; j1 < 32; ++j1) xbuf[xaddr(lane, j1)] = v[j1];
}
// before inverse pass 1 (lane = j1, v[pos])
FSC_HD void xpose_load_inv(int lane, const cplx* xbuf, cplx (&v)[32]) {
#pragma unroll
    for (int pos = 0; pos < 32; ++pos) v[pos] = xbuf[xaddr(brev5(pos), lane)];
}

// half-size variant: the transpose goes through an 8 KiB buffer of doubles, real parts first, then imaginary
// parts (same wavefront count, twice the instructions, half the shared memory)
FSC_HD void xpose_store_fwd_h(int lane, double* xb, const cplx (&v)[32], int comp) {
#pragma unroll
    for (int pos = 0; pos < 32; ++pos) xb[xaddr(brev5(pos), lane)] = comp ? v[pos].y : v[pos].x;
}
FSC_HD void xpose_load_fwd_h(int lane, const double* xb, cplx (&v)[32], int comp) {
#pragma unroll
    for (int j1 = 0; j1 < 32; ++j1) { if (comp) v[j1].y = xb[xaddr(lane, j1)]; else v[j1].x = xb[xaddr(lane, j1)]; }
}
FSC_HD void xpose_store_inv_h(int lane, double* xb, const cplx (&v)[32], int comp) {
#pragma unroll
    for (int j1 = 0; j1 < 32; ++j1) xb[xaddr(lane, j1)] = comp ? v[j1].y : v[j1].x;
}
FSC_HD void xpose_load_inv_h(int lane, const double* xb, cplx (&v)[32], int comp) {
#pragma unroll
    for (int pos = 0; pos < 32; ++pos) { if (comp) v[pos].y = xb[xaddr(brev5(pos), lane)]; else v[pos].x = xb[xaddr(brev5(pos), lane)]; }
}

// ---- accumulator arithmetic -------------------------------------------------------------
// The accumulator polynomial is stored folded: pair idx (< 1024) holds coefficients idx and
// idx + 1024, so that one 2-word access feeds one complex FFT input and X^1024 is a swap+negate.
template <typename AccT> struct acc_traits;
template <> struct acc_traits<uint64_t> {
    typedef int64_t s_t;
    static constexpr int bits = 64;
};
template <> struct acc_traits<uint32_t> {
    typedef int32_t s_t;
    static constexpr int bits = 32;
};

// coefficient pair idx of X^a * P,  a in [0, 4096)
template <typename AccT>
FSC_HD pair_t<AccT> rotated_pair(const pair_t<AccT>* poly, int idx, int a) {
    const int t0 = (idx - a) & 4095;
    const pair_t<AccT> P = poly[t0 & 1023];
    const int qd = t0 >> 10;
    pair_t<AccT> r;
    const AccT px = (qd & 1) ? P.y : P.x;
    const AccT py = (qd & 1) ? P.x : P.y;
    // qd: 0 -> (+x,+y)  1 -> (+y,-x)  2 -> (-x,-y)  3 -> (-y,+x)
    r.x = (qd >= 2) ? (AccT)(0 - px) : px;
    r.y = (qd == 1 || qd == 2) ? (AccT)(0 - py) : py;
    return r;
}

// closest signed digit of the single-level gadget decomposition with base 2^base_log
template <typename AccT>
FSC_HD double decomp_digit(AccT d, int base_log) {
    typedef typename acc_traits<AccT>::s_t s_t;
    const int sh = acc_traits<AccT>::bits - base_log;
    const AccT r = (AccT)(d + ((AccT)1 << (sh - 1)));
    return (double)(int32_t)((s_t)r >> sh);
}

// head: z[j2] = digit(X^a acc - acc) at folded index lane + 32 j2
template <typename AccT>
FSC_HD void cmux_head(int lane, const pair_t<AccT>* poly, int a, int base_log, cplx (&z)[32]) {
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        const int idx = lane + 32 * j2;
        const pair_t<AccT> R = rotated_pair<AccT>(poly, idx, a);
        const pair_t<AccT> O = poly[idx];
        z[j2].x = decomp_digit<AccT>((AccT)(R.x - O.x), base_log);
        z[j2].y = decomp_digit<AccT>((AccT)(R.y - O.y), base_log);
        FSC_SCHED_FENCE(j2, FSC_FENCE_EVERY);
    }
}

// nearest integer of v (torus scale 2^64, any magnitude < 2^100) reduced to AccT
FSC_HD uint64_t to_torus64(double v) {
    const double magic = 6755399441055744.0;                       // 1.5 * 2^52
    const double q = (v * 5.421010862427522e-20 + magic) - magic;   // rint(v / 2^64)
    const double r = fma(-q, 18446744073709551616.0, v);            // exact, |r| <= 2^63
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double2ll_rn(r);
#else
    if (r >= 9223372036854775808.0) return 0x7fffffffffffffffull;
    return (uint64_t)(int64_t)llrint(r);
#endif
}
FSC_HD uint32_t to_torus32(double v) {   // v already divided by 2^32
#if defined(__CUDA_ARCH__) && !defined(FSC_TORUS32_DADD)
    // round(v) mod 2^32 through the conversion unit (one F2I.S64.F64, a unit the blind rotation leaves idle) instead of four
    // DADDs on the FP64 pipe that bounds it: the same integer as the magic-number form below for |v| < 2^63 (|v| is 2^57 rms
    // here - 2048 x 2 digits of 2^22 times 64-bit key words, scaled by 2^-32 -, 2^63 is 60 sigma away)
    return (uint32_t)(uint64_t)__double2ll_rn(v);
#endif
    const double big = 29014219670751100192948224.0;               // 1.5 * 2^84
    const double magic = 6755399441055744.0;                       // 1.5 * 2^52
    const double r = (v + big) - big;                              // nearest multiple of 2^32
    const double w = (v - r) + magic;
    uint64_t bits;
#if defined(__CUDA_ARCH__)
    bits = (uint64_t)__double_as_longlong(w);
#else
    __builtin_memcpy(&bits, &w, 8);
#endif
    return (uint32_t)bits;
}
template <typename AccT> FSC_HD AccT to_acc(double v);
template <> FSC_HD uint64_t to_acc<uint64_t>(double v) { return to_torus64(v * (1.0 / 1024.0)); }
template <> FSC_HD uint32_t to_acc<uint32_t>(double v) { return to_torus32(v * (1.0 / 4398046511104.0)); }  // 2^-42

// conversion of an already scaled value (idft32_dit_twist applies the 1/1024 or 2^-42 factor)
template <typename AccT> FSC_HD AccT to_acc_scaled(double v);
template <> FSC_HD uint64_t to_acc_scaled<uint64_t>(double v) { return to_torus64(v); }
template <> FSC_HD uint32_t to_acc_scaled<uint32_t>(double v) { return to_torus32(v); }
// tail: acc[idx] += round(y / 1024)
template <typename AccT>
FSC_HD void cmux_tail(int lane, pair_t<AccT>* poly, const cplx (&y)[32]) {
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        const int idx = lane + 32 * j2;
        pair_t<AccT> O = poly[idx];
        O.x = (AccT)(O.x + to_acc<AccT>(y[j2].x));
        O.y = (AccT)(O.y + to_acc<AccT>(y[j2].y));
        poly[idx] = O;
        FSC_SCHED_FENCE(j2, FSC_FENCE_EVERY);
    }
}

// Fourier-domain product for k = 1, l = 1: X_q <- X_0 G[0][q] + X_1 G[1][q].
// GGSW layout per CMUX step: g[(r * 4 + (p * 2 + q)) * 32 + lane], r = register position.
struct G4 { cplx g00, g01, g10, g11; };
FSC_HD void mac_one(cplx& a, cplx& b, const G4& g) {
    const cplx x0 = a, x1 = b;
    a.x = x0.x * g.g00.x - x0.y * g.g00.y + x1.x * g.g10.x - x1.y * g.g10.y;
    a.y = x0.x * g.g00.y + x0.y * g.g00.x + x1.x * g.g10.y + x1.y * g.g10.x;
    b.x = x0.x * g.g01.x - x0.y * g.g01.y + x1.x * g.g11.x - x1.y * g.g11.y;
    b.y = x0.x * g.g01.y + x0.y * g.g01.x + x1.x * g.g11.y + x1.y * g.g11.x;
}

// modulus switch of a torus element to Z_{2N}
FSC_HD int modswitch(uint64_t x) {
    return (int)((((x >> (64 - kLogN2 - 1)) + 1) >> 1) & (2 * kN - 1));
}

// accumulator initialisation: coefficient pair idx of X^{-b} * LUT  (LUT: 2048 torus values)
template <typename AccT>
FSC_HD pair_t<AccT> lut_pair(const uint64_t* lut, int idx, int b) {
    pair_t<AccT> r;
    const int sh = 64 - acc_traits<AccT>::bits;
    {
        const int t = (idx + b) & 4095;
        const uint64_t v = (t < kN) ? lut[t] : (uint64_t)0 - lut[t - kN];
        r.x = (AccT)(v >> sh);
    }
    {
        const int t = (idx + 1024 + b) & 4095;
        const uint64_t v = (t < kN) ? lut[t] : (uint64_t)0 - lut[t - kN];
        r.y = (AccT)(v >> sh);
    }
    return r;
}

// coefficient j (< 2048) of a folded polynomial, widened to the 64-bit torus
template <typename AccT>
FSC_HD uint64_t folded_coeff(const pair_t<AccT>* poly, int j) {
    const pair_t<AccT> P = poly[j & 1023];
    const AccT v = (j >= 1024) ? P.y : P.x;
    return (uint64_t)v << (64 - acc_traits<AccT>::bits);
}

// sample extraction of coefficient 0: out[0..2048) mask, out[2048] body
template <typename AccT>
FSC_HD uint64_t extract_word(const pair_t<AccT>* mask_poly, const pair_t<AccT>* body_poly, int j) {
    if (j == kN) return folded_coeff<AccT>(body_poly, 0);
    if (j == 0) return folded_coeff<AccT>(mask_poly, 0);
    return (uint64_t)0 - folded_coeff<AccT>(mask_poly, kN - j);
}

}  // namespace fsc
