// pbs_kernel.cu — batched programmable bootstrapping for sm_100a (k = 1, N = 2048, l = 1).
//
// Kernels:
//   bsk_convert_kernel   standard-domain bootstrapping key -> Fourier domain, in the exact
//                        (register position, lane) order the blind rotation consumes
//   pbs_ring_kernel      the wide-batch kernel: CTS ciphertexts per CTA, two warps each (one per GLWE
//                        polynomial); modulus switch, accumulator init from the LUT, n CMUX steps with the
//                        Fourier GGSW streamed through a TMA-fed shared-memory ring, sample extraction of
//                        coefficient 0.  Two configurations: half-step ring + tangent-form forward passes
//                        (32-bit accumulator), chunked ring + plain passes (64-bit accumulator)
//   pbs_pair_kernel      first version (one ciphertext per CTA, key straight from L2), kept as
//                        FSC_PBS_VARIANT=pair for comparison
//   negacyclic_mul_kernel  test hook; fp64_peak_kernel: the roofline probe
// (pbs_stream_kernel.cu holds the second blind-rotation kernel, used for narrow levels)
//
// Replaces (concept): tfhe 0.10.0 programmable_bootstrap_lwe_ciphertext (Cargo.lock:482-485), the
// PBS half of shortint apply_lookup_table behind every operator in src/biguint.rs:110-248.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "pbs_core2.cuh"
#include "fsc_internal.h"
#include "tma_ring.cuh"

#ifndef FSC_TX_CONSTS
#define FSC_TX_CONSTS 0      // per-lane pass-2 constants from tensor memory: measured slower (a wait per butterfly level)
#endif
#ifndef FSC_TX_ACC
#define FSC_TX_ACC 1
#endif

namespace fsc {

__constant__ cplx c_p1[16];            // forward pass 1 / inverse pass 1 node constants (re, im), g = 32 (pbs_core.cuh lane_consts)
__constant__ cplx c_wt0[16];           // forward pass 1 in pass32 form (pbs_core2.cuh pass_const, g = 32)
struct S1PlainDev {
    __device__ __forceinline__ cplx get(int ci) const { return c_p1[ci]; }
};
struct WT0Dev {
    __device__ __forceinline__ cplx get(int ci) const { return c_wt0[ci]; }
};

void pbs_init_constants() {
    cplx h[16], w[16];
    lane_consts(kP1G, h);
    for (int ci = 0; ci < 16; ++ci) w[ci] = pass_const(ci, kP1G);
    FSC_CUDA_CHECK(cudaMemcpyToSymbol(c_p1, h, sizeof(h)));
    FSC_CUDA_CHECK(cudaMemcpyToSymbol(c_wt0, w, sizeof(w)));
}

// Negacyclic FFT of the 32 x 32 complex points held by the warp (v[j2] at lane j1), plain (re, im) constants:
// Cooley-Tukey forward, Gentleman-Sande inverse (exact reverse up to the factor 1024 that the tail removes).
// c2: the 16 per-lane pass-2 constants (lane_consts(4 lane + 1)).  Full-size (16 KiB) transpose buffer ...
template <class C2>
__device__ __forceinline__ void warp_fft_fwd_c(int lane, cplx* xbuf, const C2& c2, cplx (&v)[32]) {
    dft32_fwd(v, S1PlainDev());
    xpose_store_fwd(lane, xbuf, v);
    __syncwarp();
    xpose_load_fwd(lane, xbuf, v);
    __syncwarp();
    dft32_fwd(v, c2);
}
template <class C2>
__device__ __forceinline__ void warp_fft_inv_c(int lane, cplx* xbuf, const C2& c2, cplx (&v)[32]) {
    dft32_inv(v, c2);
    xpose_store_inv(lane, xbuf, v);
    __syncwarp();
    xpose_load_inv(lane, xbuf, v);
    __syncwarp();
    dft32_inv(v, S1PlainDev());
}
// ... or the 8 KiB half-size buffer (real parts, then imaginary parts)
template <class C2>
__device__ __forceinline__ void warp_fft_fwd_h(int lane, double* xb, const C2& c2, cplx (&v)[32]) {
    dft32_fwd(v, S1PlainDev());
    xpose_store_fwd_h(lane, xb, v, 0);
    __syncwarp();
    xpose_load_fwd_h(lane, xb, v, 0);      // v[].x: transposed real parts; v[].y: still the untransposed imaginary parts
    __syncwarp();
    xpose_store_fwd_h(lane, xb, v, 1);
    __syncwarp();
    xpose_load_fwd_h(lane, xb, v, 1);
    __syncwarp();
    dft32_fwd(v, c2);
}
template <class C2>
__device__ __forceinline__ void warp_fft_inv_h(int lane, double* xb, const C2& c2, cplx (&v)[32]) {
    dft32_inv(v, c2);
    xpose_store_inv_h(lane, xb, v, 0);
    __syncwarp();
    xpose_load_inv_h(lane, xb, v, 0);
    __syncwarp();
    xpose_store_inv_h(lane, xb, v, 1);
    __syncwarp();
    xpose_load_inv_h(lane, xb, v, 1);
    __syncwarp();
    dft32_inv(v, S1PlainDev());
}
__device__ __forceinline__ void warp_fft_fwd(int lane, cplx* xbuf, const cplx (&s2)[16], cplx (&v)[32]) {
    warp_fft_fwd_c(lane, xbuf, RegConsts(s2), v);
}
__device__ __forceinline__ void warp_fft_inv(int lane, cplx* xbuf, const cplx (&s2)[16], cplx (&v)[32]) {
    warp_fft_inv_c(lane, xbuf, RegConsts(s2), v);
}
// per-lane pass-2 constants kept in a shared-memory table [16][32] instead of 32 registers
struct SmemLaneConsts {
    const cplx* t;      // already offset by lane
    __device__ __forceinline__ cplx get(int ci) const { return t[ci * 32]; }
};

__device__ __forceinline__ cplx ldg_cplx(const cplx* p) {
    const double2 d = __ldg(reinterpret_cast<const double2*>(p));
    cplx r; r.x = d.x; r.y = d.y;
    return r;
}

// ---------------------------------------------------------------------------------------
// Bootstrapping key conversion.  grid = n * 4 polynomials, block = 32.
// in : bsk [n][p][l=1][q][2048] u64 (standard domain)      out: [n][r][g = 2p+q][lane] cplx
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) bsk_convert_kernel(const uint64_t* __restrict__ bsk, cplx* __restrict__ out) {
    __shared__ cplx xbuf[1024];
    const int lane = threadIdx.x;
    const int i = blockIdx.x >> 2, g = blockIdx.x & 3;
    const uint64_t* src = bsk + (size_t)blockIdx.x * kN;
    cplx s2[16];
    lane_consts(4 * lane + 1, s2);
    cplx v[32];
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        v[j2].x = (double)(int64_t)src[lane + 32 * j2];
        v[j2].y = (double)(int64_t)src[lane + 32 * j2 + 1024];
    }
    warp_fft_fwd(lane, xbuf, s2, v);
#pragma unroll
    for (int r = 0; r < 32; ++r) out[(((size_t)i * 32 + r) * 4 + g) * 32 + lane] = v[r];
}

// ---------------------------------------------------------------------------------------
// Blind rotation + sample extraction.  One CTA of two warps per ciphertext: warp p owns polynomial p
// of the GLWE accumulator (p = 0 mask, p = 1 body) for the whole rotation.
// shared memory: acc [2][1024] pair_t<AccT>  |  xbuf [2][1024] cplx (FFT transpose + MAC exchange)
// Per CMUX step the two warps meet twice: once to swap their Fourier-domain digit polynomials
// (every output polynomial needs both inputs) and once before the exchange buffer is reused.
// ---------------------------------------------------------------------------------------
constexpr int kGDepth = 4;      // register prefetch depth (frequency slots) of the GGSW stream

template <typename AccT>
__global__ void __launch_bounds__(64, 4) pbs_pair_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                       int n, int base_log, const uint64_t* __restrict__ luts,
                                                       const uint32_t* __restrict__ lut_idx, const __grid_constant__ OutDest out_big,
                                                       const int32_t* __restrict__ out_idx, int count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int p = threadIdx.x >> 5;
    pair_t<AccT>* acc_all = reinterpret_cast<pair_t<AccT>*>(smem_raw);
    cplx* xbuf_all = reinterpret_cast<cplx*>(smem_raw + 2 * 1024 * sizeof(pair_t<AccT>));
    pair_t<AccT>* acc = acc_all + p * 1024;
    cplx* xbuf = xbuf_all + p * 1024;
    const cplx* xother = xbuf_all + (1 - p) * 1024;
    const int c = blockIdx.x;
    if (c >= count) return;
    const uint64_t* ct = in_small + (size_t)c * (n + 1);
    const uint64_t* lut = luts + (size_t)(lut_idx ? lut_idx[c] : 0) * kN;

    cplx s2[16];
    lane_consts(4 * lane + 1, s2);

    // accumulator <- (0, X^{-b} * LUT)
    {
        const int b = modswitch(ct[n]);
#pragma unroll 4
        for (int j2 = 0; j2 < 32; ++j2) {
            const int idx = lane + 32 * j2;
            pair_t<AccT> z; z.x = 0; z.y = 0;
            acc[idx] = p ? lut_pair<AccT>(lut, idx, b) : z;
        }
    }
    __syncwarp();

    // GGSW stream of this warp: G[0][p] (g = p) and G[1][p] (g = 2 + p)
    const cplx* gbase = bsk_f + lane + p * 32;
    int a_chunk = 0;
    for (int i = 0; i < n; ++i) {
        if ((i & 31) == 0) a_chunk = (i + lane < n) ? modswitch(ct[i + lane]) : 0;
        const int a = __shfl_sync(0xffffffffu, a_chunk, i & 31);
        if (a == 0) continue;      // uniform across the CTA: both warps read the same ciphertext

        const cplx* g = gbase + (size_t)i * (32 * 4 * 32);
        cplx ga[kGDepth], gb[kGDepth];
#pragma unroll
        for (int r = 0; r < kGDepth; ++r) { ga[r] = ldg_cplx(g + (r * 4) * 32); gb[r] = ldg_cplx(g + (r * 4 + 2) * 32); }

        cplx X[32];
        cmux_head<AccT>(lane, acc, a, base_log, X);
        warp_fft_fwd(lane, xbuf, s2, X);
#pragma unroll
        for (int r = 0; r < 32; ++r) xbuf[r * 32 + lane] = X[r];
        __syncthreads();

#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const cplx o = xother[r * 32 + lane];
            const cplx g0 = ga[r % kGDepth], g1 = gb[r % kGDepth];      // G[0][p], G[1][p] at this slot
            if (r + kGDepth < 32) {
                ga[r % kGDepth] = ldg_cplx(g + ((r + kGDepth) * 4) * 32);
                gb[r % kGDepth] = ldg_cplx(g + ((r + kGDepth) * 4 + 2) * 32);
            }
            const cplx x0 = p ? o : X[r];
            const cplx x1 = p ? X[r] : o;
            X[r].x = x0.x * g0.x - x0.y * g0.y + x1.x * g1.x - x1.y * g1.y;
            X[r].y = x0.x * g0.y + x0.y * g0.x + x1.x * g1.y + x1.y * g1.x;
        }
        __syncthreads();

        warp_fft_inv(lane, xbuf, s2, X);
        cmux_tail<AccT>(lane, acc, X);
        __syncwarp();
    }
    __syncthreads();

    const size_t out = (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
    for (int j = threadIdx.x; j <= kN; j += 64) store_out_word(out_big, out + j, extract_word<AccT>(acc_all, acc_all + 1024, j));
}

// ---------------------------------------------------------------------------------------
// Ring kernel: one CTA = CTS ciphertexts (2 warps each).  The Fourier GGSW of every CMUX step is
// streamed from L2 into a shared-memory ring by 1-D bulk copies (TMA, cp.async.bulk + mbarrier
// complete_tx), so that the Fourier-domain product reads the key from shared memory instead of
// waiting on L2 latency, and one L2 read feeds all CTS ciphertexts of the CTA.
//   HS  (32-bit accumulator): NCH half-step stages of 32 KiB, non-blocking producer in warp 0 a whole
//       step ahead (tma_ring.cuh HalfProducer); own-spectrum products before the pair barrier;
//       forward passes pass32 (6-FMA tangent form), inverse pass 2 pass32_inv_gs
//   !HS (64-bit accumulator): NCH chunks of 8 KiB, one elected lane of warp 0 issues 3 chunks ahead and
//       blocks on the release of the stage it refills; plain (re, im) passes
// XH: 8 KiB half-size transpose / exchange buffer per warp (real parts, then imaginary parts).
// shared memory: acc [CTS][2][1024] pair_t<AccT> | xbuf [CTS][2][XH ? 512 : 1024] cplx | ring |
//                s2tab [16][32] cplx (per-lane pass-2 constants) | full[NCH], empty[NCH] mbarriers
// ---------------------------------------------------------------------------------------

// PP (TX only): FP64 turn-taking between the two warps of a ciphertext (same sub-partition): a warp waits for its turn
// before each 32-point pass and hands the turn over after it (two named barriers per pair, ids 5..12).  Left alone the
// two warps run every pass at the same time, each at half the pipe's rate, and then idle the pipe together during their
// transposes / head / tail; taking turns pins them one pass apart, so one's transposes overlap the other's butterflies.
template <typename AccT, int CTS, int NCH, bool XH, bool HS, bool TX, bool PP = false>
__global__ void __launch_bounds__(CTS * 64, 1) pbs_ring_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                                     int n, int base_log, const uint64_t* __restrict__ luts,
                                                                     const uint32_t* __restrict__ lut_idx, const __grid_constant__ OutDest out_big,
                                                                     const int32_t* __restrict__ out_idx, int count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pair_t<AccT>* acc_all = reinterpret_cast<pair_t<AccT>*>(smem_raw);
    constexpr int kXb = XH ? 512 : 1024;                 // transpose / exchange buffer per warp, in complex units
    cplx* xbuf_all = reinterpret_cast<cplx*>(smem_raw + (size_t)CTS * 2 * 1024 * sizeof(pair_t<AccT>));
    cplx* ring = xbuf_all + (size_t)CTS * 2 * kXb;
    cplx* s2tab = ring + (size_t)NCH * (HS ? kHalfCplx : kChunkCplx);
    uint64_t* full = reinterpret_cast<uint64_t*>(s2tab + 16 * 32);
    uint64_t* empty = full + NCH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NCH; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 2 * CTS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t tmem_base = 0;
    // TX, tensor memory columns of a ciphertext's lane quarter: [0, 256) spectra of warp p = 0 | 1 (exchange),
    // [256, 320) per-lane pass-2 constants, [320, 448) own-index accumulator pairs of warp p = 0 | 1
    constexpr int kTmemCols = 512, kTmConsts = 256, kTmAcc = 320;
    if (TX) {
        uint32_t* slot = reinterpret_cast<uint32_t*>(empty + NCH);
        if (warp == 0) tmem_alloc<kTmemCols>(slot);
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        tmem_base = *slot;
    }
    __syncthreads();

    const int total_chunks = n * kChunksPerStep;
    const bool producer = !HS && threadIdx.x == 0;
    // HS: half-step ring (tma_ring.cuh HalfProducer): 32 KB bulk copies, non-blocking producer in warp 0
    const bool hs_producer = HS && warp == 0;
    HalfProducer<NCH> hprod;
    hprod.init();
    if (hs_producer) hprod.poll(lane, bsk_f, ring, full, empty, 2 * n);
    int hs_stage = 0;
    uint32_t hs_phase = 0;
    // the producer runs kAhead chunks ahead of its own consumption; the rest of the ring (NCH - kAhead - 1 chunks)
    // is slack for ciphertexts of the CTA that trail behind
    constexpr int kAhead = NCH - 2 < 3 ? NCH - 2 : 3;
    if (producer) {
        for (int u = 0; u <= kAhead && u < total_chunks; ++u) {
            mbar_arrive_expect_tx(full + u, kChunkCplx * sizeof(cplx));
            bulk_load(ring + (size_t)u * kChunkCplx, bsk_f + (size_t)u * kChunkCplx, kChunkCplx * sizeof(cplx), full + u);
        }
    }

    // ---- consumers: warp (ct, p) owns polynomial p of ciphertext ct ----
    // TX: the two warps of a ciphertext are w and w + 4, which share a TMEM lane quarter (and an SM sub-partition)
    static_assert(!TX || (HS && CTS == 4 && sizeof(AccT) == 4), "the TMEM exchange pairs warps w and w + 4 of a 4-ciphertext CTA (32-bit accumulator)");
    const int ctl = TX ? (warp & 3) : (warp >> 1), p = TX ? (warp >> 2) : (warp & 1);
    const int c_raw = blockIdx.x * CTS + ctl;
    const bool live = c_raw < count;
    const int c = live ? c_raw : count - 1;            // padding warps shadow the last ciphertext, never store
    pair_t<AccT>* acc = acc_all + (size_t)(ctl * 2 + p) * 1024;
    cplx* xbuf = xbuf_all + (size_t)(ctl * 2 + p) * kXb;
    const cplx* xother = xbuf_all + (size_t)(ctl * 2 + (1 - p)) * kXb;
    const uint64_t* ct = in_small + (size_t)c * (n + 1);
    const uint64_t* lut = luts + (size_t)(lut_idx ? lut_idx[c] : 0) * kN;

    if (warp == 0) {      // per-lane pass-2 constants, g = 4 lane + 1
        cplx tmp[16];
        if (HS) {      // pass32 form: (re, im) for the first level, (cos, tan) below
#pragma unroll
            for (int ci = 0; ci < 16; ++ci) tmp[ci] = pass_const(ci, 4 * lane + 1);
        } else {
            lane_consts(4 * lane + 1, tmp);
        }
#pragma unroll
        for (int ci = 0; ci < 16; ++ci) s2tab[ci * 32 + lane] = tmp[ci];
    }
    __syncthreads();
    const SmemLaneConsts c2s{s2tab + lane};
    const uint32_t t_quarter = tmem_base + ((uint32_t)(ctl * 32) << 16);
    const TmemLaneConsts c2t{t_quarter + kTmConsts, {}, 0};
    if (TX && p == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cplx v4[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) v4[rr] = s2tab[(k * 4 + rr) * 32 + lane];
            tmem_st4(t_quarter + kTmConsts + 16 * k, v4);
        }
        tmem_wait_st();
    }
    const uint32_t t_acc = t_quarter + kTmAcc + (uint32_t)(p * 64);      // own-index accumulator pairs (TX)
    if (TX) {
        const int b = modswitch(ct[n]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t w[16];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = lane + 32 * (8 * k + u);
                pair_t<AccT> z; z.x = 0; z.y = 0;
                if (p) z = lut_pair<AccT>(lut, idx, b);
                acc[idx] = z;
                w[2 * u] = (uint32_t)z.x; w[2 * u + 1] = (uint32_t)z.y;
            }
            tmem_stw16(t_acc + 16 * k, w);
        }
        tmem_wait_st();
        tmem_fence_before();
        pair_barrier(1 + ctl);      // the constants staged by the p = 0 warp are visible to its partner
        tmem_fence_after();
    } else {
        const int b = modswitch(ct[n]);
#pragma unroll 4
        for (int j2 = 0; j2 < 32; ++j2) {
            const int idx = lane + 32 * j2;
            pair_t<AccT> z; z.x = 0; z.y = 0;
            acc[idx] = p ? lut_pair<AccT>(lut, idx, b) : z;
        }
    }
    __syncwarp();

    // Fourier-domain product of this warp: X_p <- X_p * G[p][p] + X_{1-p} * G[1-p][p]  (g = 2 row + col)
    const int g_own = 3 * p, g_oth = 2 - p;
    const int my_turn = 5 + 2 * ctl + p, other_turn = 5 + 2 * ctl + (1 - p);
    auto turn_wait = [&]() { if constexpr (PP) asm volatile("bar.sync %0, 64;" ::"r"(my_turn) : "memory"); };
    auto turn_pass = [&]() { if constexpr (PP) asm volatile("bar.arrive %0, 64;" ::"r"(other_turn) : "memory"); };
    if (PP && p == 1) turn_pass();      // the p = 0 warp goes first
    int a_chunk = 0;
    int t = 0;                                          // ring chunk counter
    for (int i = 0; i < n; ++i) {
        if ((i & 31) == 0) a_chunk = (i + lane < n) ? modswitch(ct[i + lane]) : 0;
        const int a = __shfl_sync(0xffffffffu, a_chunk, i & 31);
        // a == 0 is not skipped: the ring is shared by all ciphertexts of the CTA; the step is an exact no-op

        cplx X[32];
        if constexpr (TX && FSC_TX_ACC) {
            // head with the own-index pairs read from tensor memory: one shared-memory load per element (the rotated pair)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t w[16];
                tmem_ldw16(t_acc + 16 * k, w);
                tmem_wait_ld();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j2 = 8 * k + u;
                    const pair_t<AccT> R = rotated_pair<AccT>(acc, lane + 32 * j2, a);
                    X[j2].x = decomp_digit<AccT>((AccT)(R.x - (AccT)w[2 * u]), base_log);
                    X[j2].y = decomp_digit<AccT>((AccT)(R.y - (AccT)w[2 * u + 1]), base_log);
                }
            }
        } else {
            cmux_head<AccT>(lane, acc, a, base_log, X);      // (the ALU/FMA-split head of pbs_head.cuh measures 3 % slower here)
        }
        if (HS) {      // forward passes in the 6-FMA tangent form (pass32), same node constants as dft32_fwd
            double* xb = reinterpret_cast<double*>(xbuf);
            turn_wait();
            pass32(X, WT0Dev());
            turn_pass();
            xpose_store_fwd_h(lane, xb, X, 0);
            __syncwarp();
            xpose_load_fwd_h(lane, xb, X, 0);
            __syncwarp();
            xpose_store_fwd_h(lane, xb, X, 1);
            __syncwarp();
            xpose_load_fwd_h(lane, xb, X, 1);
            __syncwarp();
            turn_wait();
            if constexpr (TX && FSC_TX_CONSTS) pass32(X, c2t);
            else pass32(X, c2s);
            turn_pass();
        } else if (XH) warp_fft_fwd_h(lane, reinterpret_cast<double*>(xbuf), c2s, X);
        else warp_fft_fwd_c(lane, xbuf, c2s, X);

        if constexpr (TX) {
            // Product with the partner's spectrum handed over through tensor memory: 32 slots at once, no shared-memory
            // traffic for the exchange, one pair barrier for the hand-over and one before the buffer is reused.
            if (hs_producer) {
                while (hprod.next_h < 2 * (i + 1) && hprod.next_h < 2 * n) hprod.poll(lane, bsk_f, ring, full, empty, 2 * n);
            }
            const uint32_t t_own = tmem_base + ((uint32_t)(ctl * 32) << 16) + (uint32_t)(p * 128);
            const uint32_t t_oth = tmem_base + ((uint32_t)(ctl * 32) << 16) + (uint32_t)((1 - p) * 128);
            pair_barrier(1 + ctl);                           // the partner has read the spectrum of the previous step
            tmem_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                cplx v4[4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) v4[rr] = X[k * 4 + rr];
                tmem_st4(t_own + 16 * k, v4);
            }
            tmem_wait_st();
            tmem_fence_before();
            const int st0 = hs_stage;
            mbar_wait(full + hs_stage, hs_phase);
            if (++hs_stage == NCH) { hs_stage = 0; hs_phase ^= 1; }
            const int st1 = hs_stage;
            mbar_wait(full + hs_stage, hs_phase);
            if (++hs_stage == NCH) { hs_stage = 0; hs_phase ^= 1; }
#pragma unroll
            for (int r = 0; r < 32; ++r) {                   // own-spectrum products: hide the wait for the partner
                const cplx* g = ring + (size_t)(r < 16 ? st0 : st1) * kHalfCplx + lane;
                const cplx gw = g[((r & 15) * 4 + g_own) * 32];
                const cplx x = X[r];
                X[r].x = x.x * gw.x - x.y * gw.y;
                X[r].y = x.x * gw.y + x.y * gw.x;
            }
            pair_barrier(1 + ctl);                           // the partner's spectrum is in tensor memory
            tmem_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                cplx o4[4];
                tmem_ld4(t_oth + 16 * k, o4);
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int r = k * 4 + rr;
                    const cplx* g = ring + (size_t)(r < 16 ? st0 : st1) * kHalfCplx + lane;
                    const cplx go = g[((r & 15) * 4 + g_oth) * 32];
                    X[r].x += o4[rr].x * go.x - o4[rr].y * go.y;
                    X[r].y += o4[rr].x * go.y + o4[rr].y * go.x;
                }
            }
            tmem_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
            turn_wait();
            if constexpr (FSC_TX_CONSTS) pass32_inv_gs(X, c2t);
            else pass32_inv_gs(X, c2s);
            turn_pass();
            {
                double* xb = reinterpret_cast<double*>(xbuf);
                xpose_store_inv_h(lane, xb, X, 0);
                __syncwarp();
                xpose_load_inv_h(lane, xb, X, 0);
                __syncwarp();
                xpose_store_inv_h(lane, xb, X, 1);
                __syncwarp();
                xpose_load_inv_h(lane, xb, X, 1);
                __syncwarp();
            }
            turn_wait();
            dft32_inv(X, S1PlainDev());
            turn_pass();
            // tail: own-index pairs from tensor memory, updated values to shared memory (for the rotated reads) and back
            if constexpr (!FSC_TX_ACC) cmux_tail<AccT>(lane, acc, X);
            else
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t w[16];
                tmem_ldw16(t_acc + 16 * k, w);
                tmem_wait_ld();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j2 = 8 * k + u;
                    pair_t<AccT> O;
                    O.x = (AccT)((AccT)w[2 * u] + to_acc<AccT>(X[j2].x));
                    O.y = (AccT)((AccT)w[2 * u + 1] + to_acc<AccT>(X[j2].y));
                    acc[lane + 32 * j2] = O;
                    w[2 * u] = (uint32_t)O.x; w[2 * u + 1] = (uint32_t)O.y;
                }
                tmem_stw16(t_acc + 16 * k, w);
            }
            tmem_wait_st();
            __syncwarp();
            if (hs_producer) hprod.poll(lane, bsk_f, ring, full, empty, 2 * n);
            continue;
        }
        if constexpr (HS) {
            // Half-step product (same order as pbs_stream_kernel): the own-spectrum products run between the exchange
            // store and the pair barrier, so the wait for the partner hides behind 64 FMAs; one barrier wait and one
            // arrive per 32 KB key half.
            if (hs_producer) {
                while (hprod.next_h < 2 * (i + 1) && hprod.next_h < 2 * n) hprod.poll(lane, bsk_f, ring, full, empty, 2 * n);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int r = 0; r < 16; ++r) xbuf[r * 32 + lane] = X[half * 16 + r];
                mbar_wait(full + hs_stage, hs_phase);
                const cplx* g = ring + (size_t)hs_stage * kHalfCplx + lane;
#pragma unroll
                for (int rl = 0; rl < 16; ++rl) {
                    const int r = half * 16 + rl;
                    const cplx gw = g[(rl * 4 + g_own) * 32];
                    const cplx x = X[r];
                    X[r].x = x.x * gw.x - x.y * gw.y;
                    X[r].y = x.x * gw.y + x.y * gw.x;
                }
                pair_barrier(1 + ctl);
#pragma unroll
                for (int rl = 0; rl < 16; ++rl) {
                    const int r = half * 16 + rl;
                    const cplx o = xother[rl * 32 + lane];
                    const cplx go = g[(rl * 4 + g_oth) * 32];
                    X[r].x += o.x * go.x - o.y * go.y;
                    X[r].y += o.x * go.y + o.y * go.x;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + hs_stage);
                if (++hs_stage == NCH) { hs_stage = 0; hs_phase ^= 1; }
                if (half == 0) pair_barrier(1 + ctl);      // before the second half overwrites the exchange buffer
            }
            pass32_inv_gs(X, c2s);
            pair_barrier(1 + ctl);                         // deferred: the partner has read the second half
            {
                double* xb = reinterpret_cast<double*>(xbuf);
                xpose_store_inv_h(lane, xb, X, 0);
                __syncwarp();
                xpose_load_inv_h(lane, xb, X, 0);
                __syncwarp();
                xpose_store_inv_h(lane, xb, X, 1);
                __syncwarp();
                xpose_load_inv_h(lane, xb, X, 1);
                __syncwarp();
            }
            dft32_inv(X, S1PlainDev());
            cmux_tail<AccT>(lane, acc, X);
            __syncwarp();
            if (hs_producer) hprod.poll(lane, bsk_f, ring, full, empty, 2 * n);
            continue;
        }
        // Fourier-domain product; with the half-size buffer the spectra are swapped 16 slots at a time
#pragma unroll
        for (int half = 0; half < (XH ? 2 : 1); ++half) {
            constexpr int kSlots = XH ? 16 : 32;
#pragma unroll
            for (int r = 0; r < kSlots; ++r) xbuf[r * 32 + lane] = X[half * kSlots + r];
            pair_barrier(1 + ctl);
#pragma unroll
            for (int k = 0; k < kSlots / kChunkSlots; ++k, ++t) {
                const int stage = t % NCH;
                mbar_wait(full + stage, (uint32_t)(t / NCH) & 1);
                const cplx* g = ring + (size_t)stage * kChunkCplx + lane;
#pragma unroll
                for (int rr = 0; rr < kChunkSlots; ++rr) {
                    const int rl = k * kChunkSlots + rr, r = half * kSlots + rl;
                    const cplx o = xother[rl * 32 + lane];
                    const cplx gw = g[(rr * 4 + g_own) * 32], go = g[(rr * 4 + g_oth) * 32];
                    const cplx x = X[r];
                    X[r].x = x.x * gw.x - x.y * gw.y + o.x * go.x - o.y * go.y;
                    X[r].y = x.x * gw.y + x.y * gw.x + o.x * go.y + o.y * go.x;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                if (producer) {
                    // fetch the chunk kAhead + 1 ahead into its ring stage, once every warp has released the chunk
                    // that occupied that stage one ring revolution ago
                    const int u = t + kAhead + 1;
                    if (u < total_chunks) {
                        const int ps = u % NCH;
                        if (u >= NCH) mbar_wait(empty + ps, (uint32_t)((u - NCH) / NCH) & 1);
                        mbar_arrive_expect_tx(full + ps, kChunkCplx * sizeof(cplx));
                        bulk_load(ring + (size_t)ps * kChunkCplx, bsk_f + (size_t)u * kChunkCplx, kChunkCplx * sizeof(cplx), full + ps);
                    }
                }
            }
            pair_barrier(1 + ctl);
        }

        if (XH) warp_fft_inv_h(lane, reinterpret_cast<double*>(xbuf), c2s, X);
        else warp_fft_inv_c(lane, xbuf, c2s, X);
        cmux_tail<AccT>(lane, acc, X);
        __syncwarp();
    }
    pair_barrier(1 + ctl);

    if (live) {
        const size_t out = (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
        const pair_t<AccT>* mask = acc_all + (size_t)(ctl * 2) * 1024;
        for (int j = (p * 32 + lane); j <= kN; j += 64) store_out_word(out_big, out + j, extract_word<AccT>(mask, mask + 1024, j));
    }
    if (TX) {
        tmem_fence_before();
        __syncthreads();
        if (warp == 0) { tmem_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
    }
}

// ---------------------------------------------------------------------------------------
// Test hook: c = a (torus) * b (small integers), negacyclic, through the kernel's own FFT.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) negacyclic_mul_kernel(const uint64_t* __restrict__ a, const int64_t* __restrict__ b,
                                                             uint64_t* __restrict__ c) {
    __shared__ cplx xbuf[1024];
    const int lane = threadIdx.x;
    a += (size_t)blockIdx.x * kN; b += (size_t)blockIdx.x * kN; c += (size_t)blockIdx.x * kN;
    cplx s2[16];
    lane_consts(4 * lane + 1, s2);
    cplx va[32], vb[32];
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        va[j2].x = (double)(int64_t)a[lane + 32 * j2]; va[j2].y = (double)(int64_t)a[lane + 32 * j2 + 1024];
        vb[j2].x = (double)b[lane + 32 * j2];          vb[j2].y = (double)b[lane + 32 * j2 + 1024];
    }
    warp_fft_fwd(lane, xbuf, s2, va);
    __syncwarp();
    warp_fft_fwd(lane, xbuf, s2, vb);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const cplx x = va[r], y = vb[r];
        va[r].x = x.x * y.x - x.y * y.y; va[r].y = x.x * y.y + x.y * y.x;
    }
    __syncwarp();
    warp_fft_inv(lane, xbuf, s2, va);
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        c[lane + 32 * j2] = to_acc<uint64_t>(va[j2].x);
        c[lane + 32 * j2 + 1024] = to_acc<uint64_t>(va[j2].y);
    }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
void launch_bsk_convert(const uint64_t* bsk, void* out, int n, cudaStream_t st) {
    bsk_convert_kernel<<<n * 4, 32, 0, st>>>(bsk, reinterpret_cast<cplx*>(out));
}

template <typename AccT>
static void launch_pbs_pair_t(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                              const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    const size_t smem = 2 * 1024 * sizeof(pair_t<AccT>) + 2 * 1024 * sizeof(cplx);
    ensure_dynamic_smem(reinterpret_cast<const void*>(&pbs_pair_kernel<AccT>), smem);      // per device (the opt-in is a per-device attribute)
    pbs_pair_kernel<AccT><<<count, 64, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log, luts,
                                                    lut_idx, out_big, out_idx, count);
}

template <typename AccT, int CTS, int NCH, bool XH, bool HS, bool TX = false, bool PP = false>
static void launch_pbs_ring_t(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                              const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    const size_t smem = (size_t)CTS * 2 * 1024 * sizeof(pair_t<AccT>) + (size_t)CTS * 2 * (XH ? 512 : 1024) * sizeof(cplx) +
                        (size_t)NCH * (HS ? kHalfCplx : kChunkCplx) * sizeof(cplx) + 16 * 32 * sizeof(cplx) + 2 * NCH * sizeof(uint64_t) + 16;
    ensure_dynamic_smem(reinterpret_cast<const void*>(&pbs_ring_kernel<AccT, CTS, NCH, XH, HS, TX, PP>), smem);      // per device (the opt-in is a per-device attribute)
    const int grid = (count + CTS - 1) / CTS;
    pbs_ring_kernel<AccT, CTS, NCH, XH, HS, TX, PP><<<grid, CTS * 64, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log,
                                                                        luts, lut_idx, out_big, out_idx, count);
}

// Which blind-rotation kernel a context uses (fixed at key upload, because the Fourier key layout differs):
// FSC_PBS_VARIANT = "pair" (first version, one ciphertext per CTA) | "ring" | "stream" | "auto".  Default:
// 32-bit accumulator: auto (3) = ring kernel for wide batches (70-72 k PBS/s against 68-70 k) and stream kernel for
// levels of at most two ciphertexts per SM (5.2-5.6 ms per level against 6.9-7.0 ms), one Fourier key copy each;
// 64-bit accumulator: ring (three ciphertexts per CTA do not fit beside the stream kernel's whole-step key ring).
int pbs_variant_for(int acc_bits) {
    const char* e = getenv("FSC_PBS_VARIANT");
    if (e && e[0] == 'p') return 0;
    if (e && e[0] == 'r') return 1;
    if (e && e[0] == 's' && e[1] == 'p') return 4;      // "split": the latency kernel at every width (tests)
    if (e && e[0] == 's' && e[1] == 'o') return 5;      // "solo": one warp per ciphertext for wide batches, split / stream below
    if (e && e[0] == 's') return 2;
    if (e && e[0] == 'q') return 6;
    if (e && e[0] == 'd') return 7;      // "duo": two instruction streams per warp for wide batches, split / stream below      // "quad": four warps per ciphertext at 128 registers for wide batches, split / stream below
    if (e && e[0] == 'a') return acc_bits == 32 ? 3 : 8;
    return acc_bits == 32 ? 3 : 8;      // 8: 64-bit accumulator, ring kernel up to two ciphertexts per SM, pbs_stream_tx_kernel<2, u64> above
}

// Ring kernel configuration by batch width: wide levels pack 4 (u32 accumulator) or 3 (u64) ciphertexts per CTA
// to share every key chunk; narrow levels (at most one or two ciphertexts per SM) use 1 or 2 per CTA so that each
// ciphertext gets a less contended SM and the whole step's key fits in the ring - that is the latency-bound case
// of the carry-propagation levels.
void launch_pbs(int variant, int acc_bits, const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, int sm_count, cudaStream_t st) {
    if (count <= 0) return;
    // ring kernel, half-step key ring (NH stages of 32 KB): wide levels pack 4 (u32 accumulator) or 3 (u64) ciphertexts
    // per CTA so that one key copy feeds them all; levels of at most one or two ciphertexts per SM use 1 or 2 per CTA
#define FSC_RING(ACC, CTS, NH) \
    launch_pbs_ring_t<ACC, CTS, NH, true, true>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st)
    if (variant == 0) {      // variants 1 and 3 use the ring kernel here
        if (acc_bits == 32) launch_pbs_pair_t<uint32_t>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st);
        else launch_pbs_pair_t<uint64_t>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st);
    } else if (acc_bits == 32) {
        if (count <= sm_count) FSC_RING(uint32_t, 1, 3);
        else if (count <= 2 * sm_count) FSC_RING(uint32_t, 2, 3);
        else if (getenv("FSC_RING_NO_TMEM")) FSC_RING(uint32_t, 4, 2);      // partner exchange through shared memory (comparison)
        else if (getenv("FSC_RING_PP")) launch_pbs_ring_t<uint32_t, 4, 2, true, true, true, true>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st);
        else launch_pbs_ring_t<uint32_t, 4, 2, true, true, true>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st);
    } else {
        // 64-bit accumulator: the chunked ring (8 KB chunks, blocking producer) measures faster than the half-step ring
        // here (47.0 k against 42.1 k PBS/s at 4096 blocks)
#define FSC_RING_CHUNKED(CTS, NCH, XH) \
    launch_pbs_ring_t<uint64_t, CTS, NCH, XH, false>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st)
        if (count <= sm_count) FSC_RING_CHUNKED(1, 10, false);
        else if (count <= 2 * sm_count) FSC_RING_CHUNKED(2, 10, false);
        else FSC_RING_CHUNKED(3, 9, true);
#undef FSC_RING_CHUNKED
    }
#undef FSC_RING
}

// ---------------------------------------------------------------------------------------
// FP64 FMA peak probe (the roofline denominator north_star asks for; MEASURED_PEAKS.json has no FP64
// figure).  16 independent FMA chains per thread, each reading its own multiplier register (two fresh 64-bit operand
// pairs per DFMA: the pattern that measures highest on this part, tools/ubench/fp64_operands.cu - 36.1 TFLOP/s against
// 34.0 for chains that share both other operands), 512 FMAs per loop trip, 4 warps per SM sub-partition.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fp64_peak_kernel(double* sink, int iters, double a, double b) {
    double x[16], y[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { x[u] = threadIdx.x + u; y[u] = a + 1e-12 * (threadIdx.x + u); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 32; ++r)
#pragma unroll
            for (int u = 0; u < 16; ++u) x[u] = fma(x[u], y[u], b);
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 16; ++u) s += x[u];
    if (s == 12345.678) sink[0] = s;      // never true; keeps the chains alive
}

// returns the number of FMAs executed
double launch_fp64_peak(double* sink, int sm_count, int iters, cudaStream_t st) {
    const int blocks = sm_count * 4;      // 4 CTAs x 128 threads per SM: 4 warps per sub-partition
    fp64_peak_kernel<<<blocks, 128, 0, st>>>(sink, iters, 0.999999, 1e-9);
    return (double)blocks * 128.0 * (double)iters * 512.0;
}

void launch_negacyclic_mul(const uint64_t* a, const int64_t* b, uint64_t* c, int count, cudaStream_t st) {
    negacyclic_mul_kernel<<<count, 32, 0, st>>>(a, b, c);
}

}  // namespace fsc
