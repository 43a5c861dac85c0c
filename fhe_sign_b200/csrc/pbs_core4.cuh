// pbs_core4.cuh — per-lane building blocks of the "quad" blind rotation: the split formulation (four warps per ciphertext,
// pbs_core3.cuh) rebuilt for WIDE batches — four ciphertexts per SM, sixteen warps, 128 registers per thread.
//
// pbs_ring_kernel keeps a polynomial's 32 x 32 points in one warp (128 data registers): two warps per SM sub-partition is
// all the register file holds, and whenever both are in an integer / transpose / exchange phase the FP64 pipe idles (41 %
// of a step, profiles/README.md).  Here warp (p, h) holds 16 points per lane of polynomial p (64 data registers), so FOUR
// warps share a sub-partition — the four warps of ONE ciphertext: warps ct, ct + 4, ct + 8, ct + 12 of the CTA all own TMEM
// lane quarter ct, and everything one lane hands to the same lane of a sibling warp goes through tensor memory, not through
// the shared-memory pipe (pbs_split_kernel pays 1 270 shared-memory wavefronts per warp-step for exactly that).
//
// Slot ownership in every 32-point pass of the stream formulation (pbs_core2.cuh: same tables, transposes, key order, twist):
//   level 1 (the only butterfly level that joins the two halves)   warp h runs the eight WHOLE butterflies of slot pairs
//            (8 h + u, 16 + 8 h + u), u < 8:  v[u] = lo, v[8 + u] = hi                               (no redundant multiply)
//   join     warp 0 keeps v[0..8) (slots 0..7) and hands v[8..16) (slots 16..23) to warp 1; warp 1 keeps v[8..16)
//            (slots 24..31) and hands v[0..8) (slots 8..15) to warp 0: afterwards v[jj] is slot 16 h + jj      (8 complex each way)
//   levels 2..5 on slots [16 h, 16 h + 16)                                                           (split_levels25)
// so a pass needs the inputs of slots {8 h + u, 16 + 8 h + u}: the head computes exactly those digits, the transposed loads
// fetch exactly those columns, the Fourier-domain product forms exactly those frequencies.
// FP64 instructions per warp and CMUX step: 4 x (64 + 192) passes + 128 product + 64 twist + 128 rounding + 32 conversions
// = 1 376 (x 4 warps = 5 504 per sub-partition against 5 792 in the ring kernel).
//
// Everything here is __host__ __device__ with `lane` and `h` as arguments: tests/emu/pbs_emu4.cpp runs it on the CPU.
#pragma once
#include "pbs_core3.cuh"

namespace fsc {

// ---- transpose buffer: [32 rows][32 columns] complex, physical column = column ^ row ------------------------------
// stores (one row per instruction, column = lane) and loads (one column per instruction, row = lane or -lane mod 32) are
// both free of bank conflicts at 16 bytes per lane without a padded row: 16 KiB per polynomial exactly.
constexpr int kQuadTCplx = 32 * 32;
FSC_HD constexpr int quad_t_index(int row, int col) { return row * 32 + (col ^ row); }

// ---- level 1: eight whole butterflies, (re, im) constant (table entry 0) ----------------------------------------------
template <class SP>
FSC_HD void quad_level1(const SP& sp, cplx (&v)[16]) {
    const cplx s = sp.get(0);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const cplx lo = v[u], hi = v[8 + u];
        const double tx = fma(-s.y, hi.y, s.x * hi.x);
        const double ty = fma(s.y, hi.x, s.x * hi.y);
        v[u].x = lo.x + tx;     v[u].y = lo.y + ty;
        v[8 + u].x = lo.x - tx; v[8 + u].y = lo.y - ty;
    }
}

// ---- transposes --------------------------------------------------------------------------------------------------------
// outputs of forward pass 1 / inverse pass A: slot pos = 16 h + jj -> row brev5(pos) = brev5(jj) + h, column lane
FSC_HD void quad_xp_store(int lane, int h, cplx* T, const cplx (&w)[16]) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) T[quad_t_index(brev5(jj) + h, lane)] = w[jj];
}
// inputs of forward pass 2 / inverse pass B: row `row` (forward: lane; inverse: -lane mod 32), columns 8 h + u and 16 + 8 h + u
FSC_HD void quad_xp_load(int row, int h, const cplx* T, cplx (&v)[16]) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        v[u] = T[quad_t_index(row, 8 * h + u)];
        v[8 + u] = T[quad_t_index(row, 16 + 8 * h + u)];
    }
}

// ---- head (portable statement; the kernel runs the ALU / FMA-pipe form quad_head_u32 of pbs_quad_kernel.cu) -----------------
// digits of X^a acc - acc at folded indices lane + 32 j2 for j2 = 16 b + 8 h + u  ->  v[8 b + u]
// scratch: the polynomial's 1024 pairs by index (rotated reads); own(b, u): the own-index pair (tensor memory in the kernel)
template <typename AccT, class OWN>
FSC_HD void quad_head(int lane, int h, const pair_t<AccT>* scratch, const OWN& own, int a, int base_log, cplx (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j2 = 16 * b + 8 * h + u;
            const int idx = lane + 32 * j2;
            const pair_t<AccT> R = rotated_pair<AccT>(scratch, idx, a);
            const pair_t<AccT> O = own(b, u);
            v[8 * b + u].x = decomp_digit<AccT>((AccT)(R.x - O.x), base_log);
            v[8 * b + u].y = decomp_digit<AccT>((AccT)(R.y - O.y), base_log);
        }
}

// ---- accumulator in tensor memory -----------------------------------------------------------------------------------------
// A polynomial's own-index pairs occupy 64 columns of the lane: [parity of j2][j2 >> 1][x, y].  The tail of warp h updates
// the pairs of parity h (slot pos = 16 h + jj  <->  j2 = -brev5(pos) mod 32 = (32 - brev5(jj) - h) & 31); the head of warp h
// reads j2 = 16 b + 8 h + u: for each (b, parity) four pairs in 8 consecutive columns.
FSC_HD constexpr int quad_acc_col(int j2) { return 32 * (j2 & 1) + 2 * (j2 >> 1); }
FSC_HD constexpr int quad_tail_j2(int jj, int h) { return (32 - brev5(jj) - h) & 31; }

// ---- Fourier-domain product ------------------------------------------------------------------------------------------------
// Input slot s (frequency k1 = s) of inverse pass A is X_p[o] G[p][p] + X_{1-p}[o] G[1-p][p] with o = brev5(s) the slot that
// holds frequency s after forward pass 2.  Warp h needs s = 16 b + 8 h + u: o = 4 rev3(u) + 2 h + b — the two inputs of one
// level-1 butterfly (b = 0, 1) are NEIGHBOURING old slots, one 8-column tensor-memory load per polynomial.
FSC_HD constexpr int quad_old_slot(int u, int h) { return brev5(u) + 2 * h; }      // b = 0; b = 1 is the next slot
struct QuadKey {            // the two GGSW entries of slot s = 16 b + 8 h + u with h at run time (two compile-time candidates)
    const cplx* g0;         // key half 0 (+ lane): [16 positions][4 g][32 lanes]
    const cplx* g1;         // key half 1 (+ lane)
    int g_own, g_oth;
    FSC_HD void load(int u, int b, int h, cplx& gw, cplx& go) const {
        const int sA = u + 16 * b, sB = sA + 8;
        const int rA = freq_pos(sA), rB = freq_pos(sB);
        const cplx* gA = ((rA >> 4) ? g1 : g0) + (rA & 15) * 128;
        const cplx* gB = ((rB >> 4) ? g1 : g0) + (rB & 15) * 128;
        const cplx* g = h ? gB : gA;
        gw = g[g_own * 32]; go = g[g_oth * 32];
    }
};
FSC_HD cplx quad_mac(const cplx& x, const cplx& o, const cplx& gw, const cplx& go) {
    cplx y;
    y.x = fma(-o.y, go.y, fma(o.x, go.x, fma(-x.y, gw.y, x.x * gw.x)));
    y.y = fma(o.y, go.x, fma(o.x, go.y, fma(x.y, gw.x, x.x * gw.y)));
    return y;
}

// ---- tail: twist and rounding of the 16 outputs of inverse pass B -> 32 accumulator increments (d[2 jj], d[2 jj + 1]) ---------
FSC_HD void quad_tail_delta(int lane, int h, const cplx* tw, const cplx (&y)[16], uint32_t (&d)[32]) {
#pragma unroll
    for (int b0 = 0; b0 < 16; b0 += 8) {
        cplx t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = tw[(16 * h + b0 + u) * 32 + lane];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const cplx x = y[b0 + u];
            d[2 * (b0 + u)] = to_acc_scaled<uint32_t>(fma(-x.y, t[u].y, x.x * t[u].x));
            d[2 * (b0 + u) + 1] = to_acc_scaled<uint32_t>(fma(x.y, t[u].x, x.x * t[u].y));
        }
    }
}
// R: the 32 accumulator words of parity h in tensor-memory order (R[2 k + c] = component c of pair j2 = 2 k + h)
template <int H>
FSC_HD void quad_tail_add(const uint32_t (&d)[32], uint32_t (&R)[32]) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const int k = quad_tail_j2(jj, H) >> 1;
        R[2 * k] += d[2 * jj];
        R[2 * k + 1] += d[2 * jj + 1];
    }
}

}  // namespace fsc
