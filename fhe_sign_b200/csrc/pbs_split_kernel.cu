// pbs_split_kernel.cu — programmable bootstrapping for sm_100a (k = 1, N = 2048, l = 1), latency form for narrow levels.
//
// A PBS level of at most one ciphertext per SM (the carry-propagation levels of every radix operator: 16 to 128 blocks)
// is n sequential CMUX steps on an otherwise idle SM.  pbs_stream_kernel gives such a ciphertext two warps, one per SM
// sub-partition, and half of the SM's FP64 pipes idle; this kernel gives it four — warp (p, h) owns slots
// [16 h, 16 h + 16) of GLWE polynomial p in every pass (pbs_core3.cuh) — so every sub-partition issues half of the
// instructions per step.  Same tables, same Fourier key layout and ring as the stream kernel; everything that crosses
// warps goes through shared memory and five barriers per step (two of the block, three of a polynomial's two warps):
//
//      head (16 slots)                       -> E_p [slot][lane]          | barrier
//      pass 0 on half h                      -> T_p (transpose buffer)    | barrier
//      pass 1 on half h                      -> E_p (spectrum)            | barrier
//      product (both spectra, key ring) feeding pass 2 on half h -> T_p   | barrier
//      pass 3 on half h -> twist, rounding, accumulation (16 slots)       | barrier
//
// One CTA = one ciphertext = 128 threads.  Shared memory: acc [2][1024] pairs | E [2][32][32] complex |
// T [2][32][33] complex | key ring [NH][2048] complex (32 KB TMA bulk copies) | tables [2080] complex | mbarriers.
//
// Replaces (concept): tfhe 0.10.0 programmable_bootstrap_lwe_ciphertext (Cargo.lock:482-485), the PBS half of
// shortint apply_lookup_table behind every operator in src/biguint.rs:110-248.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "pbs_core3.cuh"
#include "fsc_internal.h"
#include "tma_ring.cuh"
#include "pbs_stream_tables.cuh"
#include "pbs_head.cuh"

namespace fsc {

// head of half h for the 32-bit accumulator in the ALU / FMA-pipe form of stream_head_u32 (pbs_head.cuh): same result
// as split_head<uint32_t>, no predicates, no conversion unit
__device__ __forceinline__ void split_head_u32(int lane, int h, const pair_t<uint32_t>* poly, int a, int base_log, cplx* E) {
    const int sh = 32 - base_log;
    const int half = 1 << (sh - 1);
    const int base = (lane - a) & 4095;
    const int q0 = base >> 10, q1 = (q0 + 1) & 3;
    const int swA = q0 & 1;
    const int sxA = 1 - (q0 & 2), syA = 1 - ((q0 ^ (q0 << 1)) & 2);
    const int sxB = 1 - (q1 & 2), syB = 1 - ((q1 ^ (q1 << 1)) & 2);
    const int dsx = sxB - sxA, dsy = syB - syA;
    const unsigned b8 = ((unsigned)(base & 1023) << 3) + 4096u * h;
    const char* pb = reinterpret_cast<const char*>(poly);
    const char* po = pb + lane * 8 + 4096 * h;
    cplx* e = E + (16 * h) * 32 + lane;
    // all 32 loads first: the warp is alone on its sub-partition, nothing else would hide their latency
    uint2 Pv[16], Ov[16];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const unsigned u = b8 + 256u * jj;                       // byte offset of the rotated pair, bit 13 = crossed
        Pv[jj] = *reinterpret_cast<const uint2*>(pb + (u & 8191u));
        Ov[jj] = *reinterpret_cast<const uint2*>(po + 256 * jj);
    }
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const unsigned u = b8 + 256u * jj;
        const int c = (int)(u >> 13);
        const uint2 P = Pv[jj], O = Ov[jj];
        const int sw = swA ^ c;
        const int sx = imad(c, dsx, sxA), sy = imad(c, dsy, syA);
        const int d = (int)(P.y - P.x);
        const int px = imad(sw, d, (int)P.x);
        const int py = (int)(P.x + P.y) - px;
        const int dx = imad(px, sx, half - (int)O.x);
        const int dy = imad(py, sy, half - (int)O.y);
        cplx z;
#ifndef FSC_HEAD_MANTISSA      // int -> double through the conversion unit (see pbs_head.cuh)
        z.x = (double)(dx >> sh);
        z.y = (double)(dy >> sh);
#else
        z.x = __hiloint2double(0x43300000, (dx >> sh) ^ (int)0x80000000) - 4503601774854144.0;
        z.y = __hiloint2double(0x43300000, (dy >> sh) ^ (int)0x80000000) - 4503601774854144.0;
#endif
        e[jj * 32] = z;
    }
}
template <typename AccT>
__device__ __forceinline__ void split_head_dev(int lane, int h, const pair_t<AccT>* poly, int a, int base_log, cplx* E) {
    if constexpr (sizeof(AccT) == 4) split_head_u32(lane, h, poly, a, base_log, E);
    else split_head<AccT>(lane, h, poly, a, base_log, E);
}
#ifdef FSC_SPLIT_GENERIC      // comparison build: every pass through the generic routines
constexpr bool kSplitUniform = false;
#else
constexpr bool kSplitUniform = true;
#endif
__device__ __forceinline__ void poly_barrier(int p) {      // the two warps of one polynomial
    if (p) asm volatile("bar.sync 2, 64;" ::: "memory");
    else asm volatile("bar.sync 1, 64;" ::: "memory");
}

template <typename AccT, int NH, int PX>
__global__ void __launch_bounds__(128, 1) pbs_split_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                            int n, int base_log, const uint64_t* __restrict__ luts,
                                                            const uint32_t* __restrict__ lut_idx, const __grid_constant__ OutDest out_big,
                                                            const int32_t* __restrict__ out_idx, int count,
                                                            const cplx* __restrict__ tabs_g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pair_t<AccT>* acc_all = reinterpret_cast<pair_t<AccT>*>(smem_raw);
    cplx* E_all = reinterpret_cast<cplx*>(smem_raw + (size_t)2 * 1024 * sizeof(pair_t<AccT>));
    cplx* T_all = E_all + 2 * kSplitECplx;
    cplx* ring = T_all + 2 * kSplitTCplx;
    cplx* tabs = ring + (size_t)NH * kHalfCplx;
    cplx* X_all = tabs + kTabCplx;                     // PX: level-1 outputs handed to the partner warp after the product
    uint64_t* full = reinterpret_cast<uint64_t*>(X_all + (PX == 2 ? 2 * kSplitX2Cplx : PX == 1 ? 2 * kSplitXCplx : 0));
    uint64_t* empty = full + NH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    static_assert(NH >= 2, "the ring must hold a whole step");

    if (threadIdx.x == 0) {
        for (int s = 0; s < NH; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < kTabCplx; t += 128) {
        const double2 d = __ldg(reinterpret_cast<const double2*>(tabs_g + t));
        tabs[t].x = d.x; tabs[t].y = d.y;
    }
    __syncthreads();

    const int total_halves = 2 * n;
    const bool producer = warp == 0;                   // warp-uniform
    HalfProducer<NH> prod;
    prod.init();
#define FSC_POLL() do { if (producer) prod.poll(lane, bsk_f, ring, full, empty, total_halves); } while (0)
    FSC_POLL();

    const int p = warp >> 1, h = warp & 1;
    const int c = blockIdx.x;                          // grid = count
    pair_t<AccT>* acc = acc_all + (size_t)p * 1024;
    cplx* E = E_all + (size_t)p * kSplitECplx;
    const cplx* E_oth = E_all + (size_t)(1 - p) * kSplitECplx;
    cplx* T = T_all + (size_t)p * kSplitTCplx;
    const uint64_t* ct = in_small + (size_t)c * (n + 1);
    const uint64_t* lut = luts + (size_t)(lut_idx ? lut_idx[c] : 0) * kN;
    {
        const int b = modswitch(ct[n]);
#pragma unroll 4
        for (int jj = 0; jj < 16; ++jj) {
            const int idx = lane + 32 * (16 * h + jj);
            pair_t<AccT> z; z.x = 0; z.y = 0;
            acc[idx] = p ? lut_pair<AccT>(lut, idx, b) : z;
        }
    }
    __syncthreads();

    const int row_inv = (32 - lane) & 31;
    const StridedConsts c0{tabs + kTabU0, 1}, c1{tabs + kTabL1 + lane, 32}, c2{tabs + kTabU2, 1}, c3{tabs + kTabL3 + lane, 32};
    int a_chunk = 0;
    int stage = 0;
    uint32_t phase = 0;
    cplx w[16];
    for (int i = 0; i < n; ++i) {
        if ((i & 31) == 0) a_chunk = (i + lane < n) ? modswitch(ct[i + lane]) : 0;
        const int a = __shfl_sync(0xffffffffu, a_chunk, i & 31);

        if constexpr (PX == 2) {
            // Rolled form (FSC_SPLIT_CFG=22, comparison): ONE copy of level 1 (strided loads), of levels 2..5 and of the
            // transpose store serves the four passes, the product runs with the half index at run time: 2 300 instructions
            // in the kernel instead of 3 350.  Measured 1-3 % slower than the unrolled form (3.88 against 3.86 ms at 148
            // blocks): the instruction cache is not what bounds four warps per SM.
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                const StridedConsts sp{tabs + (q == 0 ? kTabU0 : q == 1 ? kTabL1 + lane : q == 2 ? kTabU2 : kTabL3 + lane), (q & 1) ? 32 : 1};
                if (q == 0) {
                    split_head_dev<AccT>(lane, h, acc, a, base_log, E);
                    poly_barrier(p);
                    FSC_POLL();
                }
                if (q != 2) {
                    const SplitLoadS ld{q == 0 ? E + lane : T + (q == 1 ? lane : row_inv) * kSplitTRow, q == 0 ? 32 : 1};
                    split_level1(h, ld, sp, w);
                } else {
                    const int st0 = stage;
                    mbar_wait(full + stage, phase);
                    if (++stage == NH) { stage = 0; phase ^= 1; }
                    const int st1 = stage;
                    mbar_wait(full + stage, phase);
                    if (++stage == NH) { stage = 0; phase ^= 1; }
                    const SplitLoadProduct ld{E + lane, E_oth + lane, ring + (size_t)st0 * kHalfCplx + lane,
                                              ring + (size_t)st1 * kHalfCplx + lane, 3 * p, 2 - p};
                    cplx* X2 = X_all + (size_t)p * kSplitX2Cplx;
                    split_product_send2(lane, h, ld, sp, X2);
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
                    poly_barrier(p);
                    split_product_recv2(lane, h, X2, w);
                }
                split_levels25(h, sp, w);
                if (!(q & 1)) {
                    split_xp_store(lane, h, T, w);
                    if (q == 0) poly_barrier(p);
                    else { __syncthreads(); FSC_POLL(); }
                } else if (q == 1) {
                    split_spec_store(lane, h, E, w);
                    if (producer) {      // both halves of this step requested before any warp sleeps on them (see below)
                        while (prod.next_h < 2 * (i + 1) && prod.next_h < total_halves)
                            prod.poll(lane, bsk_f, ring, full, empty, total_halves);
                    }
                    __syncthreads();
                } else {
                    split_tail<AccT>(lane, h, acc, tabs + kTabTwist, w);
                    poly_barrier(p);
                }
            }
            continue;
        }
        // E_p and T_p are written and read by the two warps of polynomial p, except for the product, which reads both
        // spectra: block barriers on either side of the product, polynomial barriers elsewhere
        split_head_dev<AccT>(lane, h, acc, a, base_log, E);
        poly_barrier(p);
        FSC_POLL();
        if constexpr (PX == 1 && kSplitUniform) {      // uniform pass, root parameter at compile time (pbs_core3.cuh split_level1_u)
            split_level1_u<32>(h, SplitLoadE{E + lane}, c0, w);
            split_levels25(h, c0, w);
        } else {
            split_pass(h, SplitLoadE{E + lane}, c0, w);
        }
        split_xp_store(lane, h, T, w);
        poly_barrier(p);
        split_pass(h, SplitLoadT{T + lane * kSplitTRow}, c1, w);
        split_spec_store(lane, h, E, w);
        // both halves of this step must have been requested before any warp sleeps on them; the stages the producer
        // may have to wait for were released in the product of the previous step, which every warp has left
        if (producer) {
            while (prod.next_h < 2 * (i + 1) && prod.next_h < total_halves)
                prod.poll(lane, bsk_f, ring, full, empty, total_halves);
        }
        __syncthreads();
        {
            const int st0 = stage;
            mbar_wait(full + stage, phase);
            if (++stage == NH) { stage = 0; phase ^= 1; }
            const int st1 = stage;
            mbar_wait(full + stage, phase);
            if (++stage == NH) { stage = 0; phase ^= 1; }
            const SplitLoadProduct ld{E + lane, E_oth + lane, ring + (size_t)st0 * kHalfCplx + lane,
                                      ring + (size_t)st1 * kHalfCplx + lane, 3 * p, 2 - p};
            if constexpr (PX == 1) {
                cplx* X = X_all + (size_t)p * kSplitXCplx;
                split_product_send<kSplitUniform>(lane, h, ld, c2, X, w);      // g = 0: the level-1 constant is 1
                __syncwarp();
                if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
                poly_barrier(p);
                split_product_recv(lane, h, X, w);
                split_levels25_u<kSplitUniform ? 0 : 32>(h, c2, w);        // ... and so is the level-2 constant
            } else {
                split_pass(h, ld, c2, w);
                __syncwarp();
                if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
            }
        }
        split_xp_store(lane, h, T, w);
        __syncthreads();
        FSC_POLL();
        split_pass(h, SplitLoadT{T + row_inv * kSplitTRow}, c3, w);
        split_tail<AccT>(lane, h, acc, tabs + kTabTwist, w);
        poly_barrier(p);
    }
#undef FSC_POLL
    __syncthreads();

    const size_t out = (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
    for (int j = threadIdx.x; j <= kN; j += 128) store_out_word(out_big, out + j, extract_word<AccT>(acc_all, acc_all + 1024, j));
}

template <typename AccT, int NH, int PX>
static void launch_pbs_split_t(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                               const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    const size_t smem = (size_t)2 * 1024 * sizeof(pair_t<AccT>) + (size_t)2 * (kSplitECplx + kSplitTCplx) * sizeof(cplx) +
                        (size_t)NH * kHalfCplx * sizeof(cplx) + (size_t)kTabCplx * sizeof(cplx) + 2 * NH * sizeof(uint64_t) +
                        (PX == 2 ? (size_t)2 * kSplitX2Cplx : PX == 1 ? (size_t)2 * kSplitXCplx : 0) * sizeof(cplx);
    ensure_dynamic_smem(reinterpret_cast<const void*>(&pbs_split_kernel<AccT, NH, PX>), smem);      // per device (the opt-in is a per-device attribute)
    pbs_split_kernel<AccT, NH, PX><<<count, 128, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log, luts, lut_idx,
                                                         out_big, out_idx, count, stream_tables<AccT>());
}

// bsk_f: the stream kernel's Fourier key (launch_bsk_convert_stream).  One CTA per ciphertext: meant for count <= SMs.
void launch_pbs_split(int acc_bits, const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                      const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    if (count <= 0) return;
#define FSC_SPLIT(ACC, NH, PX) launch_pbs_split_t<ACC, NH, PX>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st)
    const char* cfg = getenv("FSC_SPLIT_CFG");      // comparison switch: "3r" = three-stage ring + redundant product, "2r", "22" = rolled step (measured 1-3 % slower); default: split product, unrolled
    if (acc_bits == 32) {
        if (cfg && cfg[0] == '3') FSC_SPLIT(uint32_t, 3, 0);
        else if (cfg && cfg[0] == '2' && cfg[1] == 'r') FSC_SPLIT(uint32_t, 2, 0);
        else if (cfg && cfg[0] == '2' && cfg[1] == '2') FSC_SPLIT(uint32_t, 2, 2);
        else FSC_SPLIT(uint32_t, 2, 1);
    } else {
        if (cfg && cfg[1] == 'r') FSC_SPLIT(uint64_t, 2, 0);
        else if (cfg && cfg[1] == '2') FSC_SPLIT(uint64_t, 2, 2);
        else FSC_SPLIT(uint64_t, 2, 1);
    }
#undef FSC_SPLIT
}

}  // namespace fsc
