// pbs_core3.cuh — per-lane building blocks of the "split" blind rotation: the latency form for narrow PBS levels.
//
// A level of at most one ciphertext per SM is bound by the latency of one CMUX step, not by throughput: with two
// warps per ciphertext (pbs_stream_kernel) each warp issues every instruction of a 1024-point transform alone and
// half of the SM's sub-partitions idle.  Here FOUR warps share a ciphertext: warp (p, h) owns slots [16 h, 16 h + 16)
// of polynomial p in every pass of the stream formulation (pbs_core2.cuh: same tables, same transposes, same key
// order, same twist).  pass32's first butterfly level joins slot j with slot 16 + j — the only level that crosses the
// two halves — and levels 2..5 stay inside a half, so a warp
//   * reads all 32 inputs of its lane from shared memory (the partner warp's half was written there one block
//     barrier earlier), forms its 16 level-1 outputs (lo + s hi for h = 0, lo - s hi for h = 1: 6 FP64 instructions
//     per output instead of 8 per butterfly),
//   * runs levels 2..5 on its 16 slots (the node constants of half h: entry 2^(L-2) + h 2^(L-3) + ...),
//   * writes its 16 outputs to the buffer the next pass reads.
// The Fourier-domain product feeds level 1 of inverse pass A directly: input slot s of that pass is
// X_p[brev5 s] G[p][p] + X_{1-p}[brev5 s] G[1-p][p] (both spectra from shared memory), computed by both warps of a
// polynomial — 256 redundant FP64 instructions that save one exchange and one barrier.
// FP64 instructions per warp and CMUX step: 4 x (96 + 192) passes + 256 product + 64 twist + 128 rounding = 1 600
// against 2 688, and a fourth of the shared-memory wavefronts of the other kernels matter here (one ciphertext per SM).
//
// h is a run-time value (one code path for the four warps, the loop body must not double): it enters through
// addresses, the sign of level 1 and a warp-uniform branch around the single node of level 2.
// Everything is __host__ __device__ with `lane` as an argument: tests/emu/pbs_emu3.cpp runs the same code on the CPU.
#pragma once
#include "pbs_core2.cuh"

namespace fsc {

constexpr int kSplitECplx = 32 * 32;             // exchange buffer of one polynomial: [slot][lane] complex, 16 KiB
constexpr int kSplitTRow = 33;                   // transpose buffer row (complex), padded: conflict-free both ways
constexpr int kSplitTCplx = 32 * kSplitTRow;     // [row][33] complex, 16.5 KiB

// ---- the pass on one half ------------------------------------------------------------------------
// ld(s): input slot s of the calling lane (s is a compile-time constant at every call site after unrolling)
template <class LD, class SP>
FSC_HD void split_level1(int h, const LD& ld, const SP& sp, cplx (&w)[16]) {
    // level 1, (re, im) constant: the outputs of half h only.  Loads in batches of 16 ahead of their arithmetic: a warp
    // is alone on its sub-partition here, nothing else hides the shared-memory latency
    const double sg = h ? -1.0 : 1.0;
    const cplx s = sp.get(0);
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += 8) {
        cplx lo[8], hi[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { lo[u] = ld(j0 + u); hi[u] = ld(16 + j0 + u); }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double tx = fma(-s.y, hi[u].y, s.x * hi[u].x);
            const double ty = fma(s.y, hi[u].x, s.x * hi[u].y);
            w[j0 + u].x = fma(sg, tx, lo[u].x);
            w[j0 + u].y = fma(sg, ty, lo[u].y);
        }
    }
}
// levels 2..5 on the 16 slots of half h
template <class SP> FSC_HD void split_levels35(int h, const SP& sp, cplx (&w)[16]);
template <class SP>
FSC_HD void split_levels25(int h, const SP& sp, cplx (&w)[16]) {
    {   // level 2: node m = h, constant entry 1, the odd node (h = 1) multiplies by i s
        const cplx s = sp.get(1);
        if (!h) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const cplx lo = w[j], hi = w[8 + j];
                const double qx = fma(-s.y, hi.y, hi.x);
                const double qy = fma(s.y, hi.x, hi.y);
                w[j].x = fma(s.x, qx, lo.x);      w[j].y = fma(s.x, qy, lo.y);
                w[8 + j].x = fma(-s.x, qx, lo.x); w[8 + j].y = fma(-s.x, qy, lo.y);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const cplx lo = w[j], hi = w[8 + j];
                const double qx = fma(-s.y, hi.y, hi.x);
                const double qy = fma(s.y, hi.x, hi.y);
                w[j].x = fma(-s.x, qy, lo.x);     w[j].y = fma(s.x, qx, lo.y);
                w[8 + j].x = fma(s.x, qy, lo.x);  w[8 + j].y = fma(-s.x, qx, lo.y);
            }
        }
    }
    split_levels35(h, sp, w);
}
template <class SP>
FSC_HD void split_levels35(int h, const SP& sp, cplx (&w)[16]) {
#pragma unroll
    for (int L = 3; L <= 5; ++L) {
        const int half = 16 >> (L - 1);
        const int ci0 = (1 << (L - 2)) + (h << (L - 3));      // pass32: entry 2^(L-2) + (m >> 1), m = h 2^(L-2) + mm
#pragma unroll
        for (int mm = 0; mm < (1 << (L - 2)); ++mm) {
            const int base = mm * 2 * half;
            const bool odd = mm & 1;
            const cplx s = sp.get(ci0 + (mm >> 1));
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const cplx lo = w[base + j], hi = w[base + half + j];
                const double qx = fma(-s.y, hi.y, hi.x);
                const double qy = fma(s.y, hi.x, hi.y);
                if (!odd) {
                    w[base + j].x = fma(s.x, qx, lo.x);         w[base + j].y = fma(s.x, qy, lo.y);
                    w[base + half + j].x = fma(-s.x, qx, lo.x); w[base + half + j].y = fma(-s.x, qy, lo.y);
                } else {
                    w[base + j].x = fma(-s.x, qy, lo.x);        w[base + j].y = fma(s.x, qx, lo.y);
                    w[base + half + j].x = fma(s.x, qy, lo.x);  w[base + half + j].y = fma(-s.x, qx, lo.y);
                }
            }
        }
    }
}

// ---- the two UNIFORM passes (pass 0: g = 32, pass 2: g = 0) with the root parameter at compile time ---------------------
// (the structure pass32_uniform exploits in pbs_core2.cuh, restricted to what does not depend on the half index h)
//   G = 32: the level-1 constant is exp(i pi / 4) = c (1 + i): s hi = c (hi.x - hi.y, hi.x + hi.y), 4 instructions per output;
//   G = 0 : the level-1 constant is 1 (2 instructions per output) and so is the level-2 constant (the odd node is i): four
//           additions per butterfly.
template <int G, class LD, class SP>
FSC_HD void split_level1_u(int h, const LD& ld, const SP& sp, cplx (&w)[16]) {
    static_assert(G == 0 || G == 32, "uniform tables only");
    const double sg = h ? -1.0 : 1.0;
    const double sgc = sg * sp.get(0).x;      // G = 32: +- cos(pi / 4)
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += 8) {
        cplx lo[8], hi[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { lo[u] = ld(j0 + u); hi[u] = ld(16 + j0 + u); }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (G == 0) {
                w[j0 + u].x = fma(sg, hi[u].x, lo[u].x);
                w[j0 + u].y = fma(sg, hi[u].y, lo[u].y);
            } else {
                const double qx = hi[u].x - hi[u].y, qy = hi[u].x + hi[u].y;
                w[j0 + u].x = fma(sgc, qx, lo[u].x);
                w[j0 + u].y = fma(sgc, qy, lo[u].y);
            }
        }
    }
}
template <class SP> FSC_HD void split_levels35(int h, const SP& sp, cplx (&w)[16]);
template <int G, class SP>
FSC_HD void split_levels25_u(int h, const SP& sp, cplx (&w)[16]) {
    if (G != 0) { split_levels25(h, sp, w); return; }
    if (!h) {      // level 2, constant 1
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const cplx lo = w[j], hi = w[8 + j];
            w[j].x = lo.x + hi.x;     w[j].y = lo.y + hi.y;
            w[8 + j].x = lo.x - hi.x; w[8 + j].y = lo.y - hi.y;
        }
    } else {       // the odd node: constant i
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const cplx lo = w[j], hi = w[8 + j];
            w[j].x = lo.x - hi.y;     w[j].y = lo.y + hi.x;
            w[8 + j].x = lo.x + hi.y; w[8 + j].y = lo.y - hi.x;
        }
    }
    split_levels35(h, sp, w);
}

template <class LD, class SP>
FSC_HD void split_pass(int h, const LD& ld, const SP& sp, cplx (&w)[16]) {
    split_level1(h, ld, sp, w);
    split_levels25(h, sp, w);
}

// ---- stages around the passes (warp (p, h), lane) ---------------------------------------------------
// head: digits of X^a acc - acc at folded indices lane + 32 j2, j2 in [16 h, 16 h + 16)  ->  E[j2][lane]
template <typename AccT>
FSC_HD void split_head(int lane, int h, const pair_t<AccT>* poly, int a, int base_log, cplx* E) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const int j2 = 16 * h + jj;
        const int idx = lane + 32 * j2;
        const pair_t<AccT> R = rotated_pair<AccT>(poly, idx, a);
        const pair_t<AccT> O = poly[idx];
        cplx z;
        z.x = decomp_digit<AccT>((AccT)(R.x - O.x), base_log);
        z.y = decomp_digit<AccT>((AccT)(R.y - O.y), base_log);
        E[j2 * 32 + lane] = z;
    }
}
struct SplitLoadS {          // input slot s at base[s * stride]: run-time stride, one code path for E- and T-fed passes
    const cplx* base;
    int stride;
    FSC_HD cplx operator()(int s) const { return base[s * stride]; }
};
struct SplitLoadE {          // input slot s of a pass fed by an exchange buffer
    const cplx* e;           // E + lane
    FSC_HD cplx operator()(int s) const { return e[s * 32]; }
};
struct SplitLoadT {          // input slot s of a pass fed by the transpose buffer: row `row`, column s
    const cplx* t;           // T + row * kSplitTRow
    FSC_HD cplx operator()(int s) const { return t[s]; }
};
// outputs of forward pass 1 / inverse pass A: slot pos = 16 h + jj goes to row brev5(pos), column lane
FSC_HD void split_xp_store(int lane, int h, cplx* T, const cplx (&w)[16]) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) T[(brev5(jj) + h) * kSplitTRow + lane] = w[jj];      // brev5(16 h + jj) = brev5(jj) + h
}
// outputs of forward pass 2 (the spectrum, slot pos holds frequency brev5(pos)) -> E[pos][lane]
FSC_HD void split_spec_store(int lane, int h, cplx* E, const cplx (&w)[16]) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) E[(16 * h + jj) * 32 + lane] = w[jj];
}
// input slot s (frequency k1 = s) of inverse pass A: both spectra times the GGSW column of polynomial p
struct SplitLoadProduct {
    const cplx* own;         // E_p + lane
    const cplx* oth;         // E_{1-p} + lane
    const cplx* g0;          // key half 0 (+ lane): [16 positions][4 g][32 lanes]
    const cplx* g1;          // key half 1 (+ lane)
    int g_own, g_oth;
    struct In { cplx x, o, gw, go; };
    FSC_HD In load(int s) const {
        const int r = freq_pos(s);
        const cplx* g = ((r >> 4) ? g1 : g0) + (r & 15) * 128;
        In v;
        v.x = own[brev5(s) * 32]; v.o = oth[brev5(s) * 32];
        v.gw = g[g_own * 32]; v.go = g[g_oth * 32];
        return v;
    }
    // the same for slot jj + 8 h + 16 b with h at run time (jj < 8 and b compile-time constants): the spectrum index is
    // affine in h (brev5(s + 8) = brev5(s) + 2), the key address is one of two compile-time offsets
    FSC_HD In load_rt(int jj, int b, int h) const {
        const int sA = jj + 16 * b, sB = sA + 8;
        const int rA = freq_pos(sA), rB = freq_pos(sB);
        const cplx* gA = ((rA >> 4) ? g1 : g0) + (rA & 15) * 128;
        const cplx* gB = ((rB >> 4) ? g1 : g0) + (rB & 15) * 128;
        const cplx* g = h ? gB : gA;
        const int si = (brev5(sA) + 2 * h) * 32;
        In v;
        v.x = own[si]; v.o = oth[si];
        v.gw = g[g_own * 32]; v.go = g[g_oth * 32];
        return v;
    }
    static FSC_HD cplx mul(const In& v) {
        cplx y;
        y.x = fma(-v.o.y, v.go.y, fma(v.o.x, v.go.x, fma(-v.x.y, v.gw.y, v.x.x * v.gw.x)));
        y.y = fma(v.o.y, v.go.x, fma(v.o.x, v.go.y, fma(v.x.y, v.gw.x, v.x.x * v.gw.y)));
        return y;
    }
    FSC_HD cplx operator()(int s) const { return mul(load(s)); }
};
// Product without the redundancy: warp h forms the products of input slots j and 16 + j for j in [8 h, 8 h + 8) only,
// runs their eight level-1 butterflies of inverse pass A whole, keeps the output of its own half (local index j) and
// hands the other one (same local index, other half) to its partner through X[h][j - 8 h][lane]; after a barrier
// of the polynomial's two warps split_product_recv completes w with the partner's eight outputs.
constexpr int kSplitXCplx = 2 * 8 * 32;          // exchange area of one polynomial, 8 KiB
// UNIT: the level-1 constant of inverse pass A is 1 (g = 0): the butterfly is lo +- hi
template <bool UNIT = false, class SP>
FSC_HD void split_product_send(int lane, int h, const SplitLoadProduct& ld, const SP& sp, cplx* X, cplx (&w)[16]) {
    const cplx s = sp.get(0);
    constexpr int B = 4;                      // slot pairs per batch: 8 B loads in flight ahead of the arithmetic
    if (!h) {
#pragma unroll
        for (int j0 = 0; j0 < 8; j0 += B) {
            SplitLoadProduct::In a[B], b[B];
#pragma unroll
            for (int u = 0; u < B; ++u) { a[u] = ld.load(j0 + u); b[u] = ld.load(16 + j0 + u); }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const cplx lo = SplitLoadProduct::mul(a[u]), hi = SplitLoadProduct::mul(b[u]);
                const double tx = UNIT ? hi.x : fma(-s.y, hi.y, s.x * hi.x);
                const double ty = UNIT ? hi.y : fma(s.y, hi.x, s.x * hi.y);
                w[j0 + u].x = lo.x + tx; w[j0 + u].y = lo.y + ty;
                cplx o; o.x = lo.x - tx; o.y = lo.y - ty;
                X[(j0 + u) * 32 + lane] = o;
            }
        }
    } else {
#pragma unroll
        for (int j0 = 0; j0 < 8; j0 += B) {
            SplitLoadProduct::In a[B], b[B];
#pragma unroll
            for (int u = 0; u < B; ++u) { a[u] = ld.load(8 + j0 + u); b[u] = ld.load(24 + j0 + u); }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const cplx lo = SplitLoadProduct::mul(a[u]), hi = SplitLoadProduct::mul(b[u]);
                const double tx = UNIT ? hi.x : fma(-s.y, hi.y, s.x * hi.x);
                const double ty = UNIT ? hi.y : fma(s.y, hi.x, s.x * hi.y);
                w[8 + j0 + u].x = lo.x - tx; w[8 + j0 + u].y = lo.y - ty;
                cplx o; o.x = lo.x + tx; o.y = lo.y + ty;
                X[(8 + j0 + u) * 32 + lane] = o;
            }
        }
    }
}
FSC_HD void split_product_recv(int lane, int h, const cplx* X, cplx (&w)[16]) {
    if (!h) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) w[8 + jj] = X[(8 + jj) * 32 + lane];
    } else {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) w[jj] = X[jj * 32 + lane];
    }
}
// The same with h at run time (one code path for both warps, the loop body must stay inside the instruction cache):
// every level-1 output goes through X2[half][local index][lane], the warp of half h reads its 16 back after the barrier.
constexpr int kSplitX2Cplx = 2 * 16 * 32;        // 16 KiB per polynomial
template <class SP>
FSC_HD void split_product_send2(int lane, int h, const SplitLoadProduct& ld, const SP& sp, cplx* X2) {
    const cplx s = sp.get(0);
    cplx* x = X2 + (8 * h) * 32 + lane;
#pragma unroll
    for (int j0 = 0; j0 < 8; j0 += 4) {
        SplitLoadProduct::In a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = ld.load_rt(j0 + u, 0, h); b[u] = ld.load_rt(j0 + u, 1, h); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const cplx lo = SplitLoadProduct::mul(a[u]), hi = SplitLoadProduct::mul(b[u]);
            const double tx = fma(-s.y, hi.y, s.x * hi.x);
            const double ty = fma(s.y, hi.x, s.x * hi.y);
            cplx o0, o1;
            o0.x = lo.x + tx; o0.y = lo.y + ty;
            o1.x = lo.x - tx; o1.y = lo.y - ty;
            x[(j0 + u) * 32] = o0;                 // slot 8 h + j: half 0, local index 8 h + j
            x[512 + (j0 + u) * 32] = o1;           // slot 16 + 8 h + j: half 1, same local index
        }
    }
}
FSC_HD void split_product_recv2(int lane, int h, const cplx* X2, cplx (&w)[16]) {
    const cplx* x = X2 + h * 512 + lane;
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = x[j * 32];
}
// tail: twist, rounding, accumulation of the 16 outputs of inverse pass B (slot pos <-> j2 = -brev5(pos) mod 32)
template <typename AccT>
FSC_HD void split_tail(int lane, int h, pair_t<AccT>* poly, const cplx* tw, const cplx (&y)[16]) {
#pragma unroll
    for (int b0 = 0; b0 < 16; b0 += 8) {
        cplx t[8];
        pair_t<AccT> O[8];
        double re[8], im[8];
        int idx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int pos = 16 * h + b0 + u;
            idx[u] = lane + 32 * ((32 - brev5(b0 + u) - h) & 31);      // tail_j2(16 h + jj)
            t[u] = tw[pos * 32 + lane];
            O[u] = poly[idx[u]];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const cplx x = y[b0 + u];
            re[u] = fma(-x.y, t[u].y, x.x * t[u].x);
            im[u] = fma(x.y, t[u].x, x.x * t[u].y);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            O[u].x = (AccT)(O[u].x + to_acc_scaled<AccT>(re[u]));
            O[u].y = (AccT)(O[u].y + to_acc_scaled<AccT>(im[u]));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) poly[idx[u]] = O[u];
    }
}

}  // namespace fsc
