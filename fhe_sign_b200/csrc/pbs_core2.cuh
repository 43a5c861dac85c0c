// pbs_core2.cuh — per-lane building blocks of the single-routine ("stream") blind rotation.
//
// pbs_core.cuh unrolls four different 32-point passes per CMUX step (two Cooley-Tukey forward passes, two
// Gentleman-Sande inverse passes): 3072 FP64 instructions and a 75 KB loop body, more than twice the SM's
// instruction cache, which forces every warp of an SM to run in lock step.  Here ONE pass routine serves all
// four passes; only its constant table differs (tools/fft_proto2.py is the numpy statement):
//
//   pass32(v, table(g)):  out[pos] = sum_j v[j] * (zeta^(g + 128 brev5(pos)))^j        zeta = exp(2 pi i / 4096)
//
//   forward  z_j (j = j1 + 32 j2) -> X_k = sum_j z_j zeta^(j (4k+1)),  k = k2 + 32 k1
//       pass(g = 32)        lane j1, slots j2   -> slot pos holds k2 = brev5(pos)
//       transpose                               -> lane k2, slot j1
//       pass(g = 4 k2 + 1)                      -> slot pos holds k1 = brev5(pos)
//   product  (bit reversal absorbed: the new slot s takes the old slot brev5(s))  -> slot k1
//   inverse  z_j = zeta^(-j) / 1024 * sum_k X_k (zeta^(-4j))^k
//       pass(g = 0)         lane k2, slots k1   -> slot pos holds j1 = -brev5(pos) mod 32
//       transpose (lane j1 reads row -j1 mod 32)-> lane j1, slot k2
//       pass(g = -4 j1)                         -> slot pos holds j2 = -brev5(pos) mod 32
//       twist by zeta^(-(j1 + 32 j2)) * scale   (table, scale = 1/1024 folded with the accumulator scale)
//
// Butterflies are  lo +- s * hi  with s = w (1 + i t) stored as (w, t) = (cos, tan): 6 FMAs.  The rounding error of
// this form does not grow with |t| (the product is scaled back by w), it only needs cos != 0, which holds for
// every constant of levels 2..5 of the four tables (min |cos| = 0.0046, checked by fft_proto2.py and the tests);
// level 1 reaches s = -i in one lane of the last pass and keeps the (re, im) form (8 instructions).
// FP64 instructions per warp and CMUX step: 4 x 512 (passes) + 128 (twist) + 256 (product) + 256 (rounding) = 2688.
//
// Everything is __host__ __device__ with `lane` as an argument: tests/emu/pbs_emu2.cpp runs the same code on the CPU.
#pragma once
#include <type_traits>
#include <utility>
#include "pbs_core.cuh"

namespace fsc {

// constant providers may fetch a whole butterfly level at once (begin_level); the others are asked node by node
template <class T, class = void> struct has_begin_level : std::false_type {};
template <class T> struct has_begin_level<T, std::void_t<decltype(std::declval<const T&>().begin_level(0))>> : std::true_type {};
template <class SP> FSC_HD void provider_begin_level(const SP& sp, int L) {
    if constexpr (has_begin_level<SP>::value) sp.begin_level(L);
}

// ---- constant tables --------------------------------------------------------------------
// entry ci of the table of a pass with root parameter g: ci = 0 -> (re, im); ci >= 1 -> (cos, tan)
FSC_HD cplx pass_const(int ci, int g) {
    const cplx s = twiddle4096(node_exponent(ci, g));
    cplx r;
    if (ci == 0) r = s;
    else { r.x = s.x; r.y = s.y / s.x; }
    return r;
}
// root parameter of pass q (0: forward 1, 1: forward 2, 2: inverse A, 3: inverse B) at lane `lane`
FSC_HD constexpr int pass_g(int q, int lane) { return q == 0 ? 32 : q == 1 ? 4 * lane + 1 : q == 2 ? 0 : -4 * lane; }

// consumption order of the 32 frequencies k1 in the Fourier-domain product: two halves closed under brev5
// (half 0: bit0 == bit4), the two members of a brev5 2-cycle adjacent, so that every 4-slot key chunk is closed too.
FSC_HD constexpr int freq_at(int position) {      // position -> k1
    constexpr int t[32] = {0, 2, 8, 4, 6, 12, 10, 14, 17, 19, 25, 21, 23, 29, 27, 31,
                           1, 16, 3, 24, 5, 20, 7, 28, 9, 18, 11, 26, 13, 22, 15, 30};
    return t[position];
}
FSC_HD constexpr int freq_pos(int k1) {            // k1 -> position
    constexpr int t[32] = {0, 16, 1, 18, 3, 20, 4, 22, 2, 24, 6, 26, 5, 28, 7, 30,
                           17, 8, 25, 9, 21, 11, 29, 12, 19, 10, 27, 14, 23, 13, 31, 15};
    return t[k1];
}

// ---- the pass ------------------------------------------------------------------------------
// SP::get(ci) returns table entry ci of the calling lane.
// L1TAN: table entry 0 holds (cos, tan) like the other entries and level 1 runs in the tangent form too (6 instructions per
// butterfly instead of 8).  Only for tables whose level-1 constant has cos != 0 in every lane: pass 1 (g = 4 lane + 1,
// min |cos| = 0.0245); pass 3 meets s = -i in lane 16 and keeps the (re, im) form.
template <bool L1TAN = false, class SP>
FSC_HD void pass32(cplx (&v)[32], const SP& sp) {
    {   // level 1: (re, im) constant
        provider_begin_level(sp, 1);
        const cplx s = sp.get(0);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const cplx lo = v[j], hi = v[16 + j];
            if (L1TAN) {
                const double qx = fma(-s.y, hi.y, hi.x);
                const double qy = fma(s.y, hi.x, hi.y);
                v[j].x = fma(s.x, qx, lo.x);       v[j].y = fma(s.x, qy, lo.y);
                v[16 + j].x = fma(-s.x, qx, lo.x); v[16 + j].y = fma(-s.x, qy, lo.y);
            } else {
                const double tx = fma(-s.y, hi.y, s.x * hi.x);
                const double ty = fma(s.y, hi.x, s.x * hi.y);
                v[j].x = lo.x + tx;      v[j].y = lo.y + ty;
                v[16 + j].x = lo.x - tx; v[16 + j].y = lo.y - ty;
            }
        }
    }
#pragma unroll
    for (int L = 2; L <= 5; ++L) {
        const int half = 16 >> (L - 1);
        provider_begin_level(sp, L);
#pragma unroll
        for (int m = 0; m < (1 << (L - 1)); ++m) {
            const int base = m * 2 * half;
            const bool odd = m & 1;
            const cplx s = sp.get((1 << (L - 2)) + (m >> 1));      // (w, t)
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const cplx lo = v[base + j], hi = v[base + half + j];
                const double qx = fma(-s.y, hi.y, hi.x);
                const double qy = fma(s.y, hi.x, hi.y);
                if (!odd) {
                    v[base + j].x = fma(s.x, qx, lo.x);         v[base + j].y = fma(s.x, qy, lo.y);
                    v[base + half + j].x = fma(-s.x, qx, lo.x); v[base + half + j].y = fma(-s.x, qy, lo.y);
                } else {   // node constant is i * s
                    v[base + j].x = fma(-s.x, qy, lo.x);        v[base + j].y = fma(s.x, qx, lo.y);
                    v[base + half + j].x = fma(s.x, qy, lo.x);  v[base + half + j].y = fma(-s.x, qx, lo.y);
                }
            }
        }
    }
}

// pass32 for a UNIFORM table (every lane the same constants: pass 0, g = 32, and pass 2, g = 0) with the root parameter G a
// compile-time value, so that the structure of the constants is visible to the compiler:
//   * node exponent 0 (s = 1; the odd node is i): the butterfly is four additions — all of level 1 and 2 and one constant of
//     each later level for G = 0 (a plain radix-2 DFT: 124 of 512 instructions go away);
//   * node exponent 512 (tan = 1): the two rotations are additions;
//   * level 1 of G = 32 is exponent 512: tangent form, 6 instructions instead of 8;
//   * sp.get(ci) reads a __constant__ table in the kernels: the remaining constants are constant-bank operands of the DFMAs
//     (no register operand, no shared-memory load) — the same arithmetic, in the same order, as pass32 wherever the constant
//     is not special.
template <int G, class SP>
FSC_HD void pass32_uniform(cplx (&v)[32], const SP& sp) {
    static_assert(G == 0 || G == 32, "uniform tables: g = 0 (inverse pass A) or g = 32 (forward pass 1)");
    if (G == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const cplx lo = v[j], hi = v[16 + j];
            v[j].x = lo.x + hi.x;      v[j].y = lo.y + hi.y;
            v[16 + j].x = lo.x - hi.x; v[16 + j].y = lo.y - hi.y;
        }
    } else {      // s = exp(i pi / 4) = c (1 + i)
        const double c = sp.get(0).x;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const cplx lo = v[j], hi = v[16 + j];
            const double qx = hi.x - hi.y, qy = hi.x + hi.y;
            v[j].x = fma(c, qx, lo.x);       v[j].y = fma(c, qy, lo.y);
            v[16 + j].x = fma(-c, qx, lo.x); v[16 + j].y = fma(-c, qy, lo.y);
        }
    }
#pragma unroll
    for (int L = 2; L <= 5; ++L) {
        const int half = 16 >> (L - 1);
#pragma unroll
        for (int m = 0; m < (1 << (L - 1)); ++m) {
            const int base = m * 2 * half;
            const bool odd = m & 1;
            const int ci = (1 << (L - 2)) + (m >> 1);
            const int e = node_exponent(ci, G) & 4095;
            const cplx s = sp.get(ci);      // (w, t)
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const cplx lo = v[base + j], hi = v[base + half + j];
                if (e == 0) {
                    if (!odd) {
                        v[base + j].x = lo.x + hi.x;        v[base + j].y = lo.y + hi.y;
                        v[base + half + j].x = lo.x - hi.x; v[base + half + j].y = lo.y - hi.y;
                    } else {
                        v[base + j].x = lo.x - hi.y;        v[base + j].y = lo.y + hi.x;
                        v[base + half + j].x = lo.x + hi.y; v[base + half + j].y = lo.y - hi.x;
                    }
                } else {
                    const double qx = (e == 512) ? hi.x - hi.y : fma(-s.y, hi.y, hi.x);
                    const double qy = (e == 512) ? hi.x + hi.y : fma(s.y, hi.x, hi.y);
                    if (!odd) {
                        v[base + j].x = fma(s.x, qx, lo.x);         v[base + j].y = fma(s.x, qy, lo.y);
                        v[base + half + j].x = fma(-s.x, qx, lo.x); v[base + half + j].y = fma(-s.x, qy, lo.y);
                    } else {
                        v[base + j].x = fma(-s.x, qy, lo.x);        v[base + j].y = fma(s.x, qx, lo.y);
                        v[base + half + j].x = fma(s.x, qy, lo.x);  v[base + half + j].y = fma(-s.x, qx, lo.y);
                    }
                }
            }
        }
    }
}

// Gentleman-Sande inverse of pass32 (exact reverse up to a factor 32) reading the same table: entry 0 is (re, im),
// entries >= 1 are (cos, tan) and are turned back into (re, im) with one multiply.  Used by the ring kernel, whose
// inverse keeps the merged-twist Gentleman-Sande form.
template <class SP>
FSC_HD void pass32_inv_gs(cplx (&v)[32], const SP& sp) {
#pragma unroll
    for (int L = 5; L >= 1; --L) {
        const int half = 16 >> (L - 1);
        provider_begin_level(sp, L);
#pragma unroll
        for (int m = 0; m < (1 << (L - 1)); ++m) {
            const int base = m * 2 * half;
            const int ci = (L == 1) ? 0 : ((1 << (L - 2)) + (m >> 1));
            const bool odd = (L > 1) && (m & 1);
            cplx s = sp.get(ci);
            if (ci > 0) s.y = s.x * s.y;
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const cplx u = v[base + j], w2 = v[base + half + j];
                const double dx = u.x - w2.x, dy = u.y - w2.y;
                v[base + j].x = u.x + w2.x; v[base + j].y = u.y + w2.y;
                const double ex = s.x * dx + s.y * dy;       // conj(s) * d
                const double ey = s.x * dy - s.y * dx;
                if (!odd) { v[base + half + j].x = ex; v[base + half + j].y = ey; }
                else      { v[base + half + j].x = ey; v[base + half + j].y = -ex; }   // -i * conj(s) * d
            }
        }
    }
}

struct StridedConsts {      // table entry ci at base[ci * stride] (stride 1: uniform table, 32: per-lane table)
    const cplx* base;
    int stride;
    FSC_HD cplx get(int ci) const { return base[ci * stride]; }
};

// ---- transposes through a [32][33] buffer of doubles, one component at a time -------------------------
// store: row = brev5(slot), column = lane (the same code after forward pass 1 and inverse pass A)
// load : row = `row` (forward: lane; inverse: -lane mod 32), column = slot
constexpr int kXRow = 33;                               // padded row: conflict-free column reads
constexpr int kXBufDoubles = 32 * kXRow;                // 1056 doubles = 8448 bytes per warp
FSC_HD void xp_store(int lane, double* xb, const cplx (&v)[32], int comp) {
#pragma unroll
    for (int pos = 0; pos < 32; ++pos) xb[brev5(pos) * kXRow + lane] = comp ? v[pos].y : v[pos].x;
}
FSC_HD void xp_load(int row, const double* xb, cplx (&v)[32], int comp) {
#pragma unroll
    for (int s = 0; s < 32; ++s) { if (comp) v[s].y = xb[row * kXRow + s]; else v[s].x = xb[row * kXRow + s]; }
}

// ---- Fourier-domain product -------------------------------------------------------------------
// After forward pass 2 slot s holds frequency k1 = brev5(s); inverse pass A wants slot k1.  The product absorbs the
// bit reversal: chunk R0 covers positions R0..R0+3 of the consumption order (frequencies freq_at(.), a set closed
// under brev5), reads its four old slots first and writes the four new ones.
// o: partner spectrum, gw / go: GGSW entries multiplying the own / the partner spectrum, all in position order.
template <int R0>
FSC_HD void mac_own(cplx (&X)[32], const cplx (&gw)[4]) {      // own spectrum x own GGSW entry, bit reversal absorbed
    cplx xin[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) xin[rr] = X[brev5(freq_at(R0 + rr))];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const cplx x = xin[rr];
        cplx y;
        y.x = fma(-x.y, gw[rr].y, x.x * gw[rr].x);
        y.y = fma(x.y, gw[rr].x, x.x * gw[rr].y);
        X[freq_at(R0 + rr)] = y;
    }
}
template <int R0>
FSC_HD void mac_oth(cplx (&X)[32], const cplx (&o)[4], const cplx (&go)[4]) {      // + partner spectrum x other entry
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        cplx y = X[freq_at(R0 + rr)];
        y.x = fma(-o[rr].y, go[rr].y, fma(o[rr].x, go[rr].x, y.x));
        y.y = fma(o[rr].y, go[rr].x, fma(o[rr].x, go[rr].y, y.y));
        X[freq_at(R0 + rr)] = y;
    }
}
template <int R0>
FSC_HD void mac_chunk(cplx (&X)[32], const cplx (&o)[4], const cplx (&gw)[4], const cplx (&go)[4]) {
    mac_own<R0>(X, gw);
    mac_oth<R0>(X, o, go);
}
// position (0..31) in the consumption order of the frequency held by slot s after forward pass 2
FSC_HD constexpr int slot_position(int s) { return freq_pos(brev5(s)); }

// ---- final twist + rounding + accumulation ------------------------------------------------------
// twist table: tw[pos * 32 + lane] = scale * zeta^(-(lane + 32 j2(pos))),  j2(pos) = -brev5(pos) mod 32
FSC_HD constexpr int tail_j2(int pos) { return (32 - brev5(pos)) & 31; }
template <typename AccT> FSC_HD double twist_scale();
template <> FSC_HD double twist_scale<uint64_t>() { return 1.0 / 1024.0; }
template <> FSC_HD double twist_scale<uint32_t>() { return 1.0 / 4398046511104.0; }      // 2^-42: 1/1024 and the 2^-32 accumulator scale
template <typename AccT>
FSC_HD cplx twist_const(int pos, int lane) {
    const cplx e = twiddle4096(4096 - (lane + 32 * tail_j2(pos)));
    cplx r; r.x = e.x * twist_scale<AccT>(); r.y = e.y * twist_scale<AccT>();
    return r;
}
template <typename AccT>
FSC_HD void stream_tail(int lane, pair_t<AccT>* poly, const cplx* tw, const cplx (&y)[32]) {
    // batches of 8 positions, stage by stage: eight independent twist / rounding chains in flight
#pragma unroll
    for (int b0 = 0; b0 < 32; b0 += 8) {
        cplx t[8];
        pair_t<AccT> O[8];
        double re[8], im[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { t[u] = tw[(b0 + u) * 32 + lane]; O[u] = poly[lane + 32 * tail_j2(b0 + u)]; }
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
        for (int u = 0; u < 8; ++u) poly[lane + 32 * tail_j2(b0 + u)] = O[u];
    }
}

}  // namespace fsc
