// pbs_duo_kernel.cu — wide-batch blind rotation, TWO INSTRUCTION STREAMS PER WARP (32-bit accumulator, k = 1, N = 2048, l = 1).
//
// Why: in every two-warps-per-ciphertext kernel (ring, stream) a warp owns one polynomial, so inside a CMUX step its work is one
// dependent chain — rotated difference, pass, transpose, pass, product, pass, transpose, pass, rounding — and while the warp
// transposes (shared-memory pipe) or forms digits (integer pipes) its FP64 pipe idles; the other warp of the sub-partition is
// in the same phase, and so are all eight warps of the SM (one fetch stream).  Measured on pbs_stream_tx_kernel<2>: the FP64-dense
// windows are 64 % of the step, the transposes 16 %, the head 18 % (profiles/README.md).  tools/ubench/interleave.cu: a warp that
// carries butterflies AND an independent 16-value transpose in one basic block takes 1 365 cycles for what costs 1 002 + 1 057
// apart — ptxas interleaves the two and the pipes overlap.
// Here a warp carries TWO independent half-polynomials: warp h of a ciphertext holds half h (pbs_core4.cuh: 16 points per lane)
// of polynomial 0 — stream A — and of polynomial 1 — stream B —, 2 x 64 data registers, and runs them ONE SEGMENT APART, so that
// every basic block between two barriers pairs an FP64-dense segment of one stream with the transpose / head / join segment of
// the other:
//      slot 1   A: join, levels 2-5, transpose store            B: head, level 1, join store
//      slot 2   A: transposed load, level 1, join store         B: join, levels 2-5, transpose store
//      slot 3   A: join, levels 2-5, spectrum -> TMEM           B: transposed load, level 1, join store
//      slot 4                                                   B: join, levels 2-5, spectrum -> TMEM
//      slot 5   A: product, level 1, join store
//      slot 6   A: join, levels 2-5, transpose store            B: product, level 1, join store
//      slot 7   A: transposed load, level 1, join store         B: join, levels 2-5, transpose store
//      slot 8   A: join, levels 2-5, twist / round / accumulate B: transposed load, level 1, join store
//      slot 9   A: head of the NEXT step, level 1, join store   B: join, levels 2-5, twist / round / accumulate
// (the product needs both spectra: that is where the two streams meet, slots 4 and 5 run one stream).  One barrier of the
// ciphertext's two warps (w and w + 4: one sub-partition, one TMEM lane quarter) ends every slot; everything a lane hands to the
// same lane of the sibling warp goes through tensor memory.  Layouts, ownership and arithmetic are pbs_quad_kernel's (a warp here
// is two of its warps), proven on the CPU by tests/emu/pbs_emu4.cpp:
//   TMEM  [0, 128) accumulator [p][parity][j2 >> 1][x, y] | [128, 384) spectra [p][slot] | [384, 512) level-1 join [2 p + h][8 complex]
//   smem  [ct][p] swizzled [32][32] complex transpose buffer (first 8 KB: by-index accumulator copy) | key ring 2 x 32 KB | tables
// Key layout: the stream kernel's.
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

struct DuoRole {          // one stream of a warp: half h of polynomial p
    int p;
    cplx* T;                               // the polynomial's transpose buffer
    pair_t<uint32_t>* scratch;             // = T: by-index accumulator copy
    uint32_t t_acc, t_spec_own, t_spec_oth, t_join_own, t_join_sib;
};

__device__ __forceinline__ void duo_join_st(const DuoRole& R, int h, const cplx (&v)[16]) {
    if (h) { tmem_st_c4(R.t_join_own, v); tmem_st_c4(R.t_join_own + 16, v + 4); }
    else   { tmem_st_c4(R.t_join_own, v + 8); tmem_st_c4(R.t_join_own + 16, v + 12); }
}
__device__ __forceinline__ void duo_join_ld(const DuoRole& R, int h, cplx (&v)[16]) {
    uint32_t w0[16], w1[16];
    tmem_ldw16(R.t_join_sib, w0);
    tmem_ldw16(R.t_join_sib + 16, w1);
    tmem_wait_ld();
    if (h) { words_to_c4(w0, v); words_to_c4(w1, v + 4); }
    else   { words_to_c4(w0, v + 8); words_to_c4(w1, v + 12); }
}
__device__ __forceinline__ void duo_spec_st(const DuoRole& R, int h, const cplx (&v)[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tmem_st_c4(R.t_spec_own + 64 * h + 16 * k, v + 4 * k);      // slot 16 h + jj at column 4 (16 h + jj)
}
__device__ __forceinline__ void duo_product(const DuoRole& R, int lane, int h, const cplx* g0, const cplx* g1, cplx (&v)[16]) {
    const QuadKey key{g0 + lane, g1 + lane, 3 * R.p, 2 - R.p};
#pragma unroll
    for (int u0 = 0; u0 < 8; u0 += 2) {
        uint32_t xa[2][8], xo[2][8];
        cplx gw[2][2], go[2][2];
#pragma unroll
        for (int du = 0; du < 2; ++du) {
            const uint32_t col = (uint32_t)(4 * brev5(u0 + du) + 8 * h);      // 4 columns per slot, slot = brev5(u) + 2 h (+ 1)
            tmem_ldw8(R.t_spec_own + col, xa[du]);
            tmem_ldw8(R.t_spec_oth + col, xo[du]);
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
}
__device__ __forceinline__ void duo_tail(const DuoRole& R, int lane, int h, const cplx* tw, const cplx (&v)[16]) {
    uint32_t d[32];
    quad_tail_delta(lane, h, tw, v, d);
    uint32_t W[32];
    tmem_ldw32(R.t_acc + 32 * h, W);
    tmem_wait_ld();
    if (h) quad_tail_add<1>(d, W); else quad_tail_add<0>(d, W);
    tmem_stw32(R.t_acc + 32 * h, W);
    uint2* sc = reinterpret_cast<uint2*>(R.scratch) + lane + 32 * h;
#pragma unroll
    for (int k = 0; k < 16; ++k) sc[64 * k] = make_uint2(W[2 * k], W[2 * k + 1]);      // pair index lane + 32 (2 k + h)
}

constexpr int kDuoCts = 4;

__global__ void __launch_bounds__(kDuoCts * 64, 1) pbs_duo_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                                   int n, int base_log, const uint64_t* __restrict__ luts,
                                                                   const uint32_t* __restrict__ lut_idx, const __grid_constant__ OutDest out_big,
                                                                   const int32_t* __restrict__ out_idx, int count,
                                                                   const cplx* __restrict__ tabs_g) {
    typedef uint32_t AccT;
    constexpr int NH = 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* T_all = reinterpret_cast<cplx*>(smem_raw);                       // [ct][p][32][32] complex
    cplx* ring = T_all + (size_t)kDuoCts * 2 * kQuadTCplx;
    cplx* tabs = ring + (size_t)NH * kHalfCplx;
    uint64_t* full = reinterpret_cast<uint64_t*>(tabs + kTabCplx);
    uint64_t* empty = full + NH;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NH);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NH; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kDuoCts * 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < kTabCplx; t += kDuoCts * 64) {
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

    const int ct = warp & 3, h = warp >> 2;            // warps ct and ct + 4: one sub-partition, one TMEM lane quarter
    const int c_raw = blockIdx.x * kDuoCts + ct;
    const bool live = c_raw < count;
    const int c = live ? c_raw : count - 1;            // padding warps shadow the last ciphertext, never store
    const uint64_t* ctp = in_small + (size_t)c * (n + 1);
    const uint64_t* lut = luts + (size_t)(lut_idx ? lut_idx[c] : 0) * kN;
    const uint32_t t_q = tmem_base + ((uint32_t)(ct * 32) << 16);
    DuoRole A, B;
    A.p = 0; B.p = 1;
    A.T = T_all + (size_t)(ct * 2) * kQuadTCplx; B.T = A.T + kQuadTCplx;
    A.scratch = reinterpret_cast<pair_t<AccT>*>(A.T); B.scratch = reinterpret_cast<pair_t<AccT>*>(B.T);
    A.t_acc = t_q + kQTmAcc; B.t_acc = A.t_acc + 64;
    A.t_spec_own = t_q + kQTmSpec; B.t_spec_own = A.t_spec_own + 128;
    A.t_spec_oth = B.t_spec_own; B.t_spec_oth = A.t_spec_own;
    A.t_join_own = t_q + kQTmJoin + 32 * h; A.t_join_sib = t_q + kQTmJoin + 32 * (h ^ 1);
    B.t_join_own = A.t_join_own + 64; B.t_join_sib = A.t_join_sib + 64;
    const int bar = 1 + ct;
#define FSC_SLOT_END() tmem_barrier(bar, 64)

    {   // accumulator <- (0, X^{-b} LUT): warp h writes the pairs of parity h of both polynomials, by index and to tensor memory
        const int b = modswitch(ctp[n]);
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            const DuoRole& R = p ? B : A;
            uint32_t W[32];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int idx = lane + 32 * (2 * k + h);
                pair_t<AccT> z; z.x = 0; z.y = 0;
                if (p) z = lut_pair<AccT>(lut, idx, b);
                R.scratch[idx] = z;
                W[2 * k] = z.x; W[2 * k + 1] = z.y;
            }
            tmem_stw32(R.t_acc + 32 * h, W);
        }
    }
    FSC_SLOT_END();

    const int row_inv = (32 - lane) & 31;
    const StridedConsts sp0 = pass_table(tabs, 0, lane), sp1 = pass_table(tabs, 1, lane), sp2 = pass_table(tabs, 2, lane),
                        sp3 = pass_table(tabs, 3, lane);
    const cplx* tw = tabs + kTabTwist;
    int stage = 0;
    uint32_t phase = 0;
    cplx va[16], vb[16];
    int a_chunk = (lane < n) ? modswitch(ctp[lane]) : 0;
    // prologue: A's first segment of step 0
    quad_head_u32(lane, h, A.scratch, A.t_acc, __shfl_sync(0xffffffffu, a_chunk, 0), base_log, va);
    quad_level1(sp0, va);
    duo_join_st(A, h, va);
    FSC_SLOT_END();

    for (int i = 0; i < n; ++i) {
        const int a = __shfl_sync(0xffffffffu, a_chunk, i & 31);
        if (((i + 1) & 31) == 0) a_chunk = (i + 1 + lane < n) ? modswitch(ctp[i + 1 + lane]) : 0;
        const int a_next = __shfl_sync(0xffffffffu, a_chunk, (i + 1) & 31);      // step n: a harmless extra head of stream A
        // ---- slot 1
        duo_join_ld(A, h, va);
        quad_head_u32(lane, h, B.scratch, B.t_acc, a, base_log, vb);
        split_levels25(h, sp0, va);
        quad_level1(sp0, vb);
        quad_xp_store(lane, h, A.T, va);
        duo_join_st(B, h, vb);
        FSC_SLOT_END();
        FSC_POLL();
        // ---- slot 2
        duo_join_ld(B, h, vb);
        quad_xp_load(lane, h, A.T, va);
        split_levels25(h, sp0, vb);
        quad_level1(sp1, va);
        quad_xp_store(lane, h, B.T, vb);
        duo_join_st(A, h, va);
        FSC_SLOT_END();
        // ---- slot 3
        duo_join_ld(A, h, va);
        quad_xp_load(lane, h, B.T, vb);
        split_levels25(h, sp1, va);
        quad_level1(sp1, vb);
        duo_spec_st(A, h, va);
        duo_join_st(B, h, vb);
        FSC_SLOT_END();
        // ---- slot 4
        duo_join_ld(B, h, vb);
        split_levels25(h, sp1, vb);
        duo_spec_st(B, h, vb);
        if (producer) {      // both halves of this step requested before any warp sleeps on them (see pbs_stream_kernel)
            while (prod.next_h < 2 * (i + 1) && prod.next_h < total_halves)
                prod.poll(lane, bsk_f, ring, full, empty, total_halves);
        }
        FSC_SLOT_END();
        // ---- slot 5
        const int st0 = stage;
        mbar_wait(full + stage, phase);
        if (++stage == NH) { stage = 0; phase ^= 1; }
        const int st1 = stage;
        mbar_wait(full + stage, phase);
        if (++stage == NH) { stage = 0; phase ^= 1; }
        const cplx* g0 = ring + (size_t)st0 * kHalfCplx;
        const cplx* g1 = ring + (size_t)st1 * kHalfCplx;
        duo_product(A, lane, h, g0, g1, va);
        quad_level1(sp2, va);
        duo_join_st(A, h, va);
        FSC_SLOT_END();
        // ---- slot 6
        duo_join_ld(A, h, va);
        duo_product(B, lane, h, g0, g1, vb);
        __syncwarp();
        if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
        split_levels25(h, sp2, va);
        quad_level1(sp2, vb);
        quad_xp_store(lane, h, A.T, va);
        duo_join_st(B, h, vb);
        FSC_SLOT_END();
        FSC_POLL();
        // ---- slot 7
        duo_join_ld(B, h, vb);
        quad_xp_load(row_inv, h, A.T, va);
        split_levels25(h, sp2, vb);
        quad_level1(sp3, va);
        quad_xp_store(lane, h, B.T, vb);
        duo_join_st(A, h, va);
        FSC_SLOT_END();
        // ---- slot 8
        duo_join_ld(A, h, va);
        quad_xp_load(row_inv, h, B.T, vb);
        split_levels25(h, sp3, va);
        quad_level1(sp3, vb);
        duo_tail(A, lane, h, tw, va);
        duo_join_st(B, h, vb);
        FSC_SLOT_END();
        // ---- slot 9
        duo_join_ld(B, h, vb);
        quad_head_u32(lane, h, A.scratch, A.t_acc, a_next, base_log, va);
        split_levels25(h, sp3, vb);
        quad_level1(sp0, va);
        duo_tail(B, lane, h, tw, vb);
        duo_join_st(A, h, va);
        FSC_SLOT_END();
    }
#undef FSC_POLL
#undef FSC_SLOT_END
    tmem_fence_before();
    __syncthreads();

    if (live) {
        const size_t out = (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
        for (int j = h * 32 + lane; j <= kN; j += 64) store_out_word(out_big, out + j, extract_word<AccT>(A.scratch, B.scratch, j));
    }
    if (warp == 0) { tmem_fence_after(); tmem_dealloc<kQTmCols>(tmem_base); }
}

// bsk_f: the stream kernel's Fourier key.  32-bit accumulator only; meant for count > 2 SMs (narrower levels: split / stream).
void launch_pbs_duo(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts, const uint32_t* lut_idx,
                    const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    if (count <= 0) return;
    const size_t smem = (size_t)kDuoCts * 2 * kQuadTCplx * sizeof(cplx) + (size_t)2 * kHalfCplx * sizeof(cplx) +
                        (size_t)kTabCplx * sizeof(cplx) + 2 * 2 * sizeof(uint64_t) + 16;
    ensure_dynamic_smem(reinterpret_cast<const void*>(&pbs_duo_kernel), smem);
    pbs_duo_kernel<<<(count + kDuoCts - 1) / kDuoCts, kDuoCts * 64, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log,
                                                                              luts, lut_idx, out_big, out_idx, count, stream_tables<uint32_t>());
}

}  // namespace fsc
