// pbs_quad_kernel.cu — wide-batch blind rotation, FOUR WARPS PER CIPHERTEXT at 128 registers (32-bit accumulator, k = 1,
// N = 2048, l = 1): sixteen warps per SM, four per sub-partition.
//
// Why: the ring kernel keeps a polynomial's 32 x 32 points in one warp — 128 data registers, two warps per sub-partition
// and no third; the passes and the product run at the FP64 pipe's limit, but the pipe idles through every integer /
// transpose / exchange phase (41 % of the step) because nothing else is resident to fill it.  Here warp (p, h) holds 16
// points per lane of polynomial p (pbs_core4.cuh), so a sub-partition holds the four warps of a ciphertext and every warp's
// serial chain per step is half as long.  The split kernel has the same ownership but pays for it in shared memory (every
// pass input read by both warps of a polynomial, spectra and level-1 outputs through exchange buffers: 1 270 wavefronts per
// warp-step, more than the SM's one shared-memory pipe can carry for sixteen warps).  Here the four warps of ciphertext ct
// are warps ct, ct + 4, ct + 8, ct + 12 of the CTA: ONE sub-partition, ONE tensor-memory lane quarter, and everything a lane
// hands to the same lane of a sibling warp goes through tensor memory:
//   columns [0, 128)    accumulator, own-index pairs: [p][parity of j2][j2 >> 1][x, y]  (no permanent shared-memory copy)
//   columns [128, 384)  spectra after forward pass 2: [p][slot][re, im]                (read by all four warps in the product)
//   columns [384, 512)  level-1 join: [warp r = 2 p + h][8 complex]                    (read by the other warp of the polynomial)
// Shared memory keeps only what crosses lanes — per polynomial one swizzled [32][32] complex transpose buffer whose first
// 8 KB double as the by-index copy of the accumulator for the rotated reads (written by the tail, when the buffer is free)
// — plus the key ring (2 x 32 KB TMA bulk copies per step) and the tables: 128 + 64 + 32.5 KB.
// One step, warp (p, h) (pair = the two warps of a polynomial, quad = the four warps of a ciphertext):
//      head (16 digits)                -> level 1 -> join [pair] -> levels 2-5 -> transpose store      [pair]
//      transposed load (16 columns)    -> level 1 -> join [pair] -> levels 2-5 -> spectrum -> TMEM      [quad]
//      product (16 frequencies, ring)  -> level 1 -> join [quad] -> levels 2-5 -> transpose store      [pair]
//      transposed load                 -> level 1 -> join [pair] -> levels 2-5 -> twist, round, accumulate [pair]
// Key layout: the stream kernel's (launch_bsk_convert_stream / bsk_exact_kernel).
//
// Replaces (concept): tfhe 0.10.0 programmable_bootstrap_lwe_ciphertext (Cargo.lock:482-485), the PBS half of
// shortint apply_lookup_table behind every operator in src/biguint.rs:110-248.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "pbs_core4.cuh"
#include "fsc_internal.h"
#include "tma_ring.cuh"
#include "pbs_stream_tables.cuh"
#include "pbs_head.cuh"
#include "pbs_quad_dev.cuh"

namespace fsc {

__global__ void __launch_bounds__(kQuadCts * 128, 1) pbs_quad_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                                     int n, int base_log, const uint64_t* __restrict__ luts,
                                                                     const uint32_t* __restrict__ lut_idx, const __grid_constant__ OutDest out_big,
                                                                     const int32_t* __restrict__ out_idx, int count,
                                                                     const cplx* __restrict__ tabs_g) {
    typedef uint32_t AccT;
    constexpr int NH = 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* T_all = reinterpret_cast<cplx*>(smem_raw);                       // [ct][p][32][32] complex
    cplx* ring = T_all + (size_t)kQuadCts * 2 * kQuadTCplx;
    cplx* tabs = ring + (size_t)NH * kHalfCplx;
    uint64_t* full = reinterpret_cast<uint64_t*>(tabs + kTabCplx);
    uint64_t* empty = full + NH;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NH);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NH; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kQuadCts * 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < kTabCplx; t += kQuadCts * 128) {
        const double2 d = __ldg(reinterpret_cast<const double2*>(tabs_g + t));
        tabs[t].x = d.x; tabs[t].y = d.y;
    }
    if (warp == 0) tmem_alloc<kQTmCols>(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_halves = 2 * n;
    const bool producer = warp == 0;                   // warp-uniform
    HalfProducer<NH> prod;
    prod.init();
#define FSC_POLL() do { if (producer) prod.poll(lane, bsk_f, ring, full, empty, total_halves); } while (0)
    FSC_POLL();

    const int ct = warp & 3, r = warp >> 2, p = r >> 1, h = r & 1;      // the four warps of a ciphertext share sub-partition ct
    const int c_raw = blockIdx.x * kQuadCts + ct;
    const bool live = c_raw < count;
    const int c = live ? c_raw : count - 1;            // padding warps shadow the last ciphertext, never store
    cplx* T = T_all + (size_t)(ct * 2 + p) * kQuadTCplx;
    pair_t<AccT>* scratch = reinterpret_cast<pair_t<AccT>*>(T);      // by-index copy of the accumulator polynomial (first 8 KB)
    const uint64_t* ctp = in_small + (size_t)c * (n + 1);
    const uint64_t* lut = luts + (size_t)(lut_idx ? lut_idx[c] : 0) * kN;
    const uint32_t t_q = tmem_base + ((uint32_t)(ct * 32) << 16);
    const uint32_t t_acc = t_q + kQTmAcc + 64 * p;
    const uint32_t t_spec_own = t_q + kQTmSpec + 128 * p, t_spec_oth = t_q + kQTmSpec + 128 * (1 - p);
    const uint32_t t_join_own = t_q + kQTmJoin + 32 * r, t_join_sib = t_q + kQTmJoin + 32 * (r ^ 1);
    const int bar_pair = 1 + 2 * ct + p, bar_quad = 9 + ct;

    {   // accumulator <- (0, X^{-b} LUT): warp (p, h) writes the pairs of parity h, by index and to tensor memory
        const int b = modswitch(ctp[n]);
        uint32_t R[32];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int idx = lane + 32 * (2 * k + h);
            pair_t<AccT> z; z.x = 0; z.y = 0;
            if (p) z = lut_pair<AccT>(lut, idx, b);
            scratch[idx] = z;
            R[2 * k] = z.x; R[2 * k + 1] = z.y;
        }
        tmem_stw32(t_acc + 32 * h, R);
    }
    tmem_barrier(bar_pair, 64);

    const int row_inv = (32 - lane) & 31;
    int a_chunk = 0;
    int stage = 0;
    uint32_t phase = 0;
    cplx v[16];
    for (int i = 0; i < n; ++i) {
        if ((i & 31) == 0) a_chunk = (i + lane < n) ? modswitch(ctp[i + lane]) : 0;
        const int a = __shfl_sync(0xffffffffu, a_chunk, i & 31);

#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const StridedConsts sp = pass_table(tabs, q, lane);
            // ---- the 16 inputs of this warp's eight level-1 butterflies: slots 8 h + u -> v[u], 16 + 8 h + u -> v[8 + u]
            if (q == 0) {
                quad_head_u32(lane, h, scratch, t_acc, a, base_log, v);
                FSC_POLL();
            } else if (q == 2) {
                if (producer) {      // both halves of this step requested before any warp sleeps on them (see pbs_stream_kernel)
                    while (prod.next_h < 2 * (i + 1) && prod.next_h < total_halves)
                        prod.poll(lane, bsk_f, ring, full, empty, total_halves);
                }
                const int st0 = stage;
                mbar_wait(full + stage, phase);
                if (++stage == NH) { stage = 0; phase ^= 1; }
                const int st1 = stage;
                mbar_wait(full + stage, phase);
                if (++stage == NH) { stage = 0; phase ^= 1; }
                const QuadKey key{ring + (size_t)st0 * kHalfCplx + lane, ring + (size_t)st1 * kHalfCplx + lane, 3 * p, 2 - p};
#pragma unroll
                for (int u0 = 0; u0 < 8; u0 += 2) {
                    uint32_t xa[2][8], xo[2][8];
                    cplx gw[2][2], go[2][2];
#pragma unroll
                    for (int du = 0; du < 2; ++du) {
                        const uint32_t col = (uint32_t)(4 * brev5(u0 + du) + 8 * h);      // 4 columns per slot, slot = brev5(u) + 2 h (+ 1)
                        tmem_ldw8(t_spec_own + col, xa[du]);
                        tmem_ldw8(t_spec_oth + col, xo[du]);
                        key.load(u0 + du, 0, h, gw[du][0], go[du][0]);
                        key.load(u0 + du, 1, h, gw[du][1], go[du][1]);
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int du = 0; du < 2; ++du)
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            const cplx x = cplx_from_words(xa[du][4 * b], xa[du][4 * b + 1], xa[du][4 * b + 2], xa[du][4 * b + 3]);
                            const cplx o = cplx_from_words(xo[du][4 * b], xo[du][4 * b + 1], xo[du][4 * b + 2], xo[du][4 * b + 3]);
                            v[8 * b + u0 + du] = quad_mac(x, o, gw[du][b], go[du][b]);
                        }
                }
                __syncwarp();
                if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
            } else {
                quad_xp_load(q == 1 ? lane : row_inv, h, T, v);
            }
            // ---- level 1 and the join: afterwards v[jj] is slot 16 h + jj
            quad_level1(sp, v);
            if (h) { tmem_st_c4(t_join_own, v); tmem_st_c4(t_join_own + 16, v + 4); }
            else   { tmem_st_c4(t_join_own, v + 8); tmem_st_c4(t_join_own + 16, v + 12); }
            tmem_barrier(q == 2 ? bar_quad : bar_pair, q == 2 ? 128 : 64);      // q == 2: every warp of the ciphertext has read the spectra
            {
                uint32_t w0[16], w1[16];
                tmem_ldw16(t_join_sib, w0);
                tmem_ldw16(t_join_sib + 16, w1);
                tmem_wait_ld();
                if (h) { words_to_c4(w0, v); words_to_c4(w1, v + 4); }
                else   { words_to_c4(w0, v + 8); words_to_c4(w1, v + 12); }
            }
            split_levels25(h, sp, v);
            // ---- what follows the pass
            if (!(q & 1)) {
                quad_xp_store(lane, h, T, v);
                named_barrier(bar_pair, 64);
                FSC_POLL();
            } else if (q == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tmem_st_c4(t_spec_own + 64 * h + 16 * k, v + 4 * k);      // slot 16 h + jj at column 4 (16 h + jj)
                tmem_barrier(bar_quad, 128);
            } else {
                uint32_t d[32];
                quad_tail_delta(lane, h, tabs + kTabTwist, v, d);
                uint32_t R[32];
                tmem_ldw32(t_acc + 32 * h, R);
                tmem_wait_ld();
                if (h) quad_tail_add<1>(d, R); else quad_tail_add<0>(d, R);
                tmem_stw32(t_acc + 32 * h, R);
                // the transpose buffer is free (both warps of the polynomial passed the join after their transposed loads):
                // its first 8 KB take the by-index copy for the next head's rotated reads
                uint2* sc = reinterpret_cast<uint2*>(scratch) + lane + 32 * h;
#pragma unroll
                for (int k = 0; k < 16; ++k) sc[64 * k] = make_uint2(R[2 * k], R[2 * k + 1]);      // pair index lane + 32 (2 k + h)
                tmem_barrier(bar_pair, 64);
            }
        }
    }
#undef FSC_POLL
    tmem_fence_before();
    __syncthreads();

    if (live) {
        const size_t out = (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
        const pair_t<AccT>* mask = reinterpret_cast<const pair_t<AccT>*>(T_all + (size_t)(ct * 2) * kQuadTCplx);
        const pair_t<AccT>* body = reinterpret_cast<const pair_t<AccT>*>(T_all + (size_t)(ct * 2 + 1) * kQuadTCplx);
        for (int j = r * 32 + lane; j <= kN; j += 128) store_out_word(out_big, out + j, extract_word<AccT>(mask, body, j));
    }
    if (warp == 0) { tmem_fence_after(); tmem_dealloc<kQTmCols>(tmem_base); }
}

// bsk_f: the stream kernel's Fourier key.  32-bit accumulator only; meant for count > 2 SMs (narrower levels: split / stream).
void launch_pbs_quad(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts, const uint32_t* lut_idx,
                     const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    if (count <= 0) return;
    const size_t smem = (size_t)kQuadCts * 2 * kQuadTCplx * sizeof(cplx) + (size_t)2 * kHalfCplx * sizeof(cplx) +
                        (size_t)kTabCplx * sizeof(cplx) + 2 * 2 * sizeof(uint64_t) + 16;
    ensure_dynamic_smem(reinterpret_cast<const void*>(&pbs_quad_kernel), smem);
    pbs_quad_kernel<<<(count + kQuadCts - 1) / kQuadCts, kQuadCts * 128, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log,
                                                                                  luts, lut_idx, out_big, out_idx, count, stream_tables<uint32_t>());
}

}  // namespace fsc
