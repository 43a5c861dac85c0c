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

// Tangent ("FMA") form of a unit constant s = exp(2 pi i e / 4096): form 1 keeps (w, t) = (cos, sin/cos) when the
// cosine dominates, form 2 keeps (sin, cos/sin) otherwise, so that |t| <= 1 and  a +- s*b  costs 6 FMAs:
//   form 1:  s*b = w * ((bx - t by) + i (by + t bx))        form 2:  s*b = w * ((t bx - by) + i (t by + bx))
FSC_HD constexpr int tan_form(int e) { return ((e & 2047) <= 512 || (e & 2047) >= 1536) ? 1 : 2; }
// form of stored constant ci of a pass whose lanes are centred on g_center; levels below min_level stay (re, im)
FSC_HD constexpr int node_form(int ci, int g_center, int min_level) {
    return node_level(ci) < min_level ? 0 : tan_form(node_exponent(ci, g_center));
}

FSC_HD void lane_consts(int g, cplx (&s)[16]) {
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) s[ci] = twiddle4096(node_exponent(ci, g));
}
FSC_HD cplx tan_const(int e, int form) {
    const cplx s = twiddle4096(e);
    cplx r;
    if (form == 1) { r.x = s.x; r.y = s.y / s.x; } else if (form == 2) { r.x = s.y; r.y = s.x / s.y; } else r = s;
    return r;
}
FSC_HD void lane_consts_tan(int g, int g_center, int min_level, cplx (&s)[16]) {
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) s[ci] = tan_const(node_exponent(ci, g), node_form(ci, g_center, min_level));
}
constexpr int kP1Center = 32, kP1MinLevel = 1;      // pass 1: uniform constants, all levels in tangent form
constexpr int kP2Center = 63, kP2MinLevel = 3;      // pass 2: per-lane constants, levels 3..5 stay inside one octant

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

// forward pass with constants stored in tangent form where node_form says so (6 FMAs per butterfly)
template <int G_CENTER, int MIN_LEVEL, class SP>
FSC_HD void dft32_fwd_tan(cplx (&v)[32], const SP& sp) {
#pragma unroll
    for (int L = 1; L <= 5; ++L) {
        const int half = 16 >> (L - 1);
#pragma unroll
        for (int m = 0; m < (1 << (L - 1)); ++m) {
            const int base = m * 2 * half;
            const int ci = (L == 1) ? 0 : ((1 << (L - 2)) + (m >> 1));
            const bool odd = (L > 1) && (m & 1);
            const int form = node_form(ci, G_CENTER, MIN_LEVEL);
            const cplx s = sp.get(ci);
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const cplx lo = v[base + j], hi = v[base + half + j];
                double tx, ty;
                if (form == 0) { tx = s.x * hi.x - s.y * hi.y; ty = s.x * hi.y + s.y * hi.x; }
                else if (form == 1) { tx = fma(-s.y, hi.y, hi.x); ty = fma(s.y, hi.x, hi.y); }
                else { tx = fma(s.y, hi.x, -hi.y); ty = fma(s.y, hi.y, hi.x); }
                const double w = form == 0 ? 1.0 : s.x;
                if (!odd) {
                    v[base + j].x = fma(w, tx, lo.x);         v[base + j].y = fma(w, ty, lo.y);
                    v[base + half + j].x = fma(-w, tx, lo.x); v[base + half + j].y = fma(-w, ty, lo.y);
                } else {   // constant is i*s
                    v[base + j].x = fma(-w, ty, lo.x);        v[base + j].y = fma(w, tx, lo.y);
                    v[base + half + j].x = fma(w, ty, lo.x);  v[base + half + j].y = fma(-w, tx, lo.y);
                }
            }
        }
    }
}

// Inverse of pass 1 as a plain inverse DFT (decimation in time: bit-reversed in, natural out, twiddle before
// add, mostly trivial twiddles) followed by the fused untwist * scale:
//   out[j2] = tw[j2] * sum_k2 v[brev5(k2)] omega_32^(-j2 k2),   tw[j2] = scale * exp(-2 pi i j2 / 128).
// cp.get(16 + j): omega_32^(-j) in tangent form, j < 16;  cp.get(tw_base + j2): tw[j2] as (re, im).
template <int ST, class SP>
FSC_HD void idft32_dit_stage(cplx (&v)[32], const SP& cp) {
    constexpr int half = 1 << (ST - 1), m = 2 * half, nblk = 32 / m;
#pragma unroll
    for (int blk = 0; blk < nblk; ++blk) {
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const int b = blk * m;
            const int widx = j * (32 / m);                     // omega_m^-j = omega_32^-(j 32/m)
            const cplx u = v[b + j], h = v[b + j + half];
            if (widx == 0) {
                v[b + j].x = u.x + h.x; v[b + j].y = u.y + h.y;
                v[b + j + half].x = u.x - h.x; v[b + j + half].y = u.y - h.y;
            } else if (widx == 8) {                            // w = -i:  w*h = (h.y, -h.x)
                v[b + j].x = u.x + h.y; v[b + j].y = u.y - h.x;
                v[b + j + half].x = u.x - h.y; v[b + j + half].y = u.y + h.x;
            } else {
                const int form = tan_form(4096 - 128 * widx);
                const cplx s = cp.get(16 + widx);
                double tx, ty;
                if (form == 1) { tx = fma(-s.y, h.y, h.x); ty = fma(s.y, h.x, h.y); }
                else { tx = fma(s.y, h.x, -h.y); ty = fma(s.y, h.y, h.x); }
                v[b + j].x = fma(s.x, tx, u.x);         v[b + j].y = fma(s.x, ty, u.y);
                v[b + j + half].x = fma(-s.x, tx, u.x); v[b + j + half].y = fma(-s.x, ty, u.y);
            }
        }
    }
}
template <class SP>
FSC_HD void idft32_dit_twist(cplx (&v)[32], const SP& cp, int tw_base) {
    idft32_dit_stage<1>(v, cp);
    idft32_dit_stage<2>(v, cp);
    idft32_dit_stage<3>(v, cp);
    idft32_dit_stage<4>(v, cp);
    idft32_dit_stage<5>(v, cp);
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        const cplx t = cp.get(tw_base + j2), x = v[j2];
        v[j2].x = x.x * t.x - x.y * t.y;
        v[j2].y = x.x * t.y + x.y * t.x;
    }
}

// table behind the uniform-constant provider of the kernels: 16 pass-1 forward constants (tangent form),
// 16 omega_32^-j (tangent form), 32 untwist constants scaled for the 64-bit accumulator, 32 for the 32-bit one
constexpr int kUniTw64 = 32, kUniTw32 = 64, kUniPlainP1 = 96, kUniSize = 112;
inline void fill_uniform_table(cplx* t) {
    cplx p1[16];
    lane_consts_tan(kP1Center, kP1Center, kP1MinLevel, p1);
    for (int i = 0; i < 16; ++i) t[i] = p1[i];
    for (int j = 0; j < 16; ++j) t[16 + j] = tan_const(4096 - 128 * j, j == 0 ? 0 : tan_form(4096 - 128 * j));
    {
        cplx pl[16];
        lane_consts(kP1Center, pl);
        for (int i = 0; i < 16; ++i) t[kUniPlainP1 + i] = pl[i];
    }
    for (int j = 0; j < 32; ++j) {
        const cplx e = twiddle4096(4096 - 32 * j);
        t[kUniTw64 + j].x = e.x * (1.0 / 1024.0);          t[kUniTw64 + j].y = e.y * (1.0 / 1024.0);
        t[kUniTw32 + j].x = e.x * (1.0 / 4398046511104.0); t[kUniTw32 + j].y = e.y * (1.0 / 4398046511104.0);
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

// inverse pass whose constants come from a tangent-form table (plain (re, im) rebuilt with one multiply)
template <int G_CENTER, int MIN_LEVEL, class SP>
FSC_HD void dft32_inv_tan(cplx (&v)[32], const SP& sp) {
#pragma unroll
    for (int L = 5; L >= 1; --L) {
        const int half = 16 >> (L - 1);
#pragma unroll
        for (int m = 0; m < (1 << (L - 1)); ++m) {
            const int base = m * 2 * half;
            const int ci = (L == 1) ? 0 : ((1 << (L - 2)) + (m >> 1));
            const bool odd = (L > 1) && (m & 1);
            const int form = node_form(ci, G_CENTER, MIN_LEVEL);
            cplx s = sp.get(ci);
            if (form == 1) { s.y = s.x * s.y; } else if (form == 2) { const double w = s.x; s.x = w * s.y; s.y = w; }
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const cplx u = v[base + j], w2 = v[base + half + j];
                const double dx = u.x - w2.x, dy = u.y - w2.y;
                v[base + j].x = u.x + w2.x; v[base + j].y = u.y + w2.y;
                const double ex = s.x * dx + s.y * dy;
                const double ey = s.x * dy - s.y * dx;
                if (!odd) { v[base + half + j].x = ex; v[base + half + j].y = ey; }
                else      { v[base + half + j].x = ey; v[base + half + j].y = -ex; }
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
template <typename AccT>
FSC_HD void cmux_tail_scaled(int lane, pair_t<AccT>* poly, const cplx (&y)[32]) {
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        const int idx = lane + 32 * j2;
        pair_t<AccT> O = poly[idx];
        O.x = (AccT)(O.x + to_acc_scaled<AccT>(y[j2].x));
        O.y = (AccT)(O.y + to_acc_scaled<AccT>(y[j2].y));
        poly[idx] = O;
        FSC_SCHED_FENCE(j2, FSC_FENCE_EVERY);
    }
}
template <typename AccT> struct uni_tw { };
template <> struct uni_tw<uint64_t> { static constexpr int base = kUniTw64; };
template <> struct uni_tw<uint32_t> { static constexpr int base = kUniTw32; };

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
