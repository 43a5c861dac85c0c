// pbs_stream_kernel.cu — batched programmable bootstrapping for sm_100a (k = 1, N = 2048, l = 1), "stream" form.
//
// Same algorithm and data flow as pbs_ring_kernel (two warps per ciphertext, warp p owns GLWE polynomial p; the
// Fourier GGSW of every CMUX step arrives in a shared-memory ring by TMA bulk copies), rebuilt around ONE 32-point
// pass routine (pbs_core2.cuh): a CMUX step is a four-trip loop
//
//      q = 0: head            -> pass(table 0) -> transpose store
//      q = 1: transpose load  -> pass(table 1) -> Fourier-domain product (key ring, partner exchange)
//      q = 2:                    pass(table 2) -> transpose store
//      q = 3: transpose load  -> pass(table 3) -> twist, rounding, accumulation
//
// so the loop body holds one copy of the butterflies instead of four: about 2 000 instructions (32 KB) against
// 4 700 (75 KB), 2 688 FP64 instructions per warp-step against 3 072 (tangent-form butterflies on every pass).
// Kernels:
//   bsk_convert_stream_kernel   standard-domain key -> Fourier domain in the product's consumption order
//   pbs_stream_kernel<AccT, CTS, NH>       the rolled form above: levels of up to two ciphertexts per SM
//   pbs_stream_tx_kernel<MODE, AccT>       wide batches, four ciphertexts per SM, tensor memory for everything a lane hands to the
//                                          same lane (partner spectra, accumulator, twist constants).  MODE 2 — the default wide-batch
//                                          kernel of the library, 91 k PBS/s at 32 bits, 76 k at 64 — is a straight-line step with the
//                                          two uniform passes specialised (pass32_uniform), whole-complex transposes through a buffer
//                                          that also holds the by-index accumulator copy, and the int <-> double conversions on the
//                                          conversion unit; MODE 0 / 1 are the rolled forms kept for comparison (FSC_STREAM_TX).
//
// Replaces (concept): tfhe 0.10.0 programmable_bootstrap_lwe_ciphertext (Cargo.lock:482-485), the PBS half of
// shortint apply_lookup_table behind every operator in src/biguint.rs:110-248.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <mutex>
#include <vector>
#include "pbs_core2.cuh"
#include "fsc_internal.h"
#include "tma_ring.cuh"
#include "pbs_head.cuh"
#include "pbs_stream_tables.cuh"

namespace fsc {

// Uniform pass tables (pass 0: g = 32, pass 2: g = 0) in constant memory for pass32_uniform: the constants become
// constant-bank operands of the DFMAs.  Written once per device by stream_uniform_constants_init().
__constant__ cplx c_su[2][16];
template <int Q> struct UniConsts {
    __device__ __forceinline__ cplx get(int ci) const { return c_su[Q][ci]; }
};
static void stream_uniform_constants_init() {
    static bool done[64] = {};
    static std::mutex mu;
    int dev = 0;
    FSC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev & 63]) return;
    cplx h[2][16];
    for (int ci = 0; ci < 16; ++ci) { h[0][ci] = pass_const(ci, pass_g(0, 0)); h[1][ci] = pass_const(ci, pass_g(2, 0)); }
    FSC_CUDA_CHECK(cudaMemcpyToSymbol(c_su, h, sizeof(h)));
    done[dev & 63] = true;
}

// forward FFT of the 32 x 32 complex points held by the warp through the shared pass (tables in global memory)
__device__ __forceinline__ void stream_fft_fwd(int lane, double* xb, const cplx* tabs, cplx (&v)[32]) {
    pass32(v, pass_table(tabs, 0, lane));
    xp_store(lane, xb, v, 0);
    __syncwarp();
    xp_load(lane, xb, v, 0);
    __syncwarp();
    xp_store(lane, xb, v, 1);
    __syncwarp();
    xp_load(lane, xb, v, 1);
    __syncwarp();
    pass32(v, pass_table(tabs, 1, lane));
}

// ---------------------------------------------------------------------------------------
// Bootstrapping key conversion.  grid = n * 4 polynomials, block = 32.
// in : bsk [n][p][l=1][q][2048] u64 (standard domain)      out: [n][position][g = 2p+q][lane] cplx
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) bsk_convert_stream_kernel(const uint64_t* __restrict__ bsk, cplx* __restrict__ out,
                                                                const cplx* __restrict__ tabs) {
    __shared__ double xb[kXBufDoubles];
    const int lane = threadIdx.x;
    const int i = blockIdx.x >> 2, g = blockIdx.x & 3;
    const uint64_t* src = bsk + (size_t)blockIdx.x * kN;
    cplx v[32];
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        v[j2].x = (double)(int64_t)src[lane + 32 * j2];
        v[j2].y = (double)(int64_t)src[lane + 32 * j2 + 1024];
    }
    stream_fft_fwd(lane, xb, tabs, v);
#pragma unroll
    for (int s = 0; s < 32; ++s) out[(((size_t)i * 32 + slot_position(s)) * 4 + g) * 32 + lane] = v[s];
}

// ---------------------------------------------------------------------------------------
// Blind rotation + sample extraction.  One CTA = CTS ciphertexts, two warps each.
// shared memory: acc [CTS][2][1024] pair_t<AccT> | xbuf [CTS][2][1056] double | ring [NH][2048] cplx |
//                tables [2080] cplx | full[NH], empty[NH] mbarriers
// ---------------------------------------------------------------------------------------
template <typename AccT, int CTS, int NH>
__global__ void __launch_bounds__(CTS * 64, 1) pbs_stream_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                                  int n, int base_log, const uint64_t* __restrict__ luts,
                                                                  const uint32_t* __restrict__ lut_idx, const __grid_constant__ OutDest out_big,
                                                                  const int32_t* __restrict__ out_idx, int count,
                                                                  const cplx* __restrict__ tabs_g, int stagger) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pair_t<AccT>* acc_all = reinterpret_cast<pair_t<AccT>*>(smem_raw);
    double* xbuf_all = reinterpret_cast<double*>(smem_raw + (size_t)CTS * 2 * 1024 * sizeof(pair_t<AccT>));
    cplx* ring = reinterpret_cast<cplx*>(xbuf_all + (size_t)CTS * 2 * kXBufDoubles);
    cplx* tabs = ring + (size_t)NH * kHalfCplx;
    uint64_t* full = reinterpret_cast<uint64_t*>(tabs + kTabCplx);
    uint64_t* empty = full + NH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    static_assert(NH >= 2, "the ring must hold a whole step");

    if (threadIdx.x == 0) {
        for (int s = 0; s < NH; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 2 * CTS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < kTabCplx; t += CTS * 64) {
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

    const int ctl = warp >> 1, p = warp & 1;
    const int c_raw = blockIdx.x * CTS + ctl;
    const bool live = c_raw < count;
    const int c = live ? c_raw : count - 1;            // padding warps shadow the last ciphertext, never store
    pair_t<AccT>* acc = acc_all + (size_t)(ctl * 2 + p) * 1024;
    double* xb = xbuf_all + (size_t)(ctl * 2 + p) * kXBufDoubles;
    cplx* xc = reinterpret_cast<cplx*>(xb);                                                            // exchange view: [16][32] cplx
    const cplx* xother = reinterpret_cast<const cplx*>(xbuf_all + (size_t)(ctl * 2 + (1 - p)) * kXBufDoubles);
    const uint64_t* ct = in_small + (size_t)c * (n + 1);
    const uint64_t* lut = luts + (size_t)(lut_idx ? lut_idx[c] : 0) * kN;
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
    // Fourier-domain product of this warp: X_p <- X_p * G[p][p] + X_{1-p} * G[1-p][p]  (g = 2 row + col)
    const int g_own = 3 * p, g_oth = 2 - p;
    const int row_inv = (32 - lane) & 31;
    int a_chunk = 0;
    int stage = 0;
    uint32_t phase = 0;
    cplx X[32];
    if (stagger > 0 && CTS == 4 && ctl >= 2) {      // experiment (FSC_STREAM_STAGGER): here warps w and w + 4 — one sub-partition — belong to
        const long long t0 = clock64();              // DIFFERENT ciphertexts (ctl = warp >> 1): the second pair of ciphertexts starts late
        while (clock64() - t0 < (long long)stagger) { }
        __syncwarp();
    }
    for (int i = 0; i < n; ++i) {
        if ((i & 31) == 0) a_chunk = (i + lane < n) ? modswitch(ct[i + lane]) : 0;
        const int a = __shfl_sync(0xffffffffu, a_chunk, i & 31);

        auto do_head = [&]() {
            stream_head<AccT>(lane, acc, a, base_log, X);
        };
        auto xp_out = [&](bool after_product) {
            if (after_product) pair_barrier(1 + ctl);      // deferred from the product: the partner has read this buffer
            xp_store(lane, xb, X, 0);
            __syncwarp();
        };
        auto xp_in = [&](int row) {
            xp_load(row, xb, X, 0);
            __syncwarp();
            xp_store(lane, xb, X, 1);
            __syncwarp();
            xp_load(row, xb, X, 1);
            __syncwarp();
        };
        auto do_mac = [&]() {
            // Both halves of this step must have been requested before any warp may sleep on them.  The producer warp
            // makes sure of that here, once per step; it cannot deadlock: the stages it waits for are released by
            // the other warps in the product of the previous step, which never waits for warp 0.
            if (producer) {
                while (prod.next_h < 2 * (i + 1) && prod.next_h < total_halves)
                    prod.poll(lane, bsk_f, ring, full, empty, total_halves);
            }
            // One half = 16 frequencies.  The own-spectrum products need no partner data: they run between the
            // exchange store and the pair barrier, so the wait for the partner warp hides behind 64 FMAs.
            auto half = [&](auto hc) {
                constexpr int H = decltype(hc)::value;
#pragma unroll
                for (int s = 0; s < 32; ++s)
                    if ((slot_position(s) >> 4) == H) xc[(slot_position(s) & 15) * 32 + lane] = X[s];
                mbar_wait(full + stage, phase);
                const cplx* g = ring + (size_t)stage * kHalfCplx + lane;
                auto own = [&](auto kc) {
                    constexpr int K = decltype(kc)::value;
                    cplx gw[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) gw[rr] = g[((K * 4 + rr) * 4 + g_own) * 32];
                    mac_own<H * 16 + K * 4>(X, gw);
                };
                own(std::integral_constant<int, 0>{});
                own(std::integral_constant<int, 1>{});
                own(std::integral_constant<int, 2>{});
                own(std::integral_constant<int, 3>{});
                pair_barrier(1 + ctl);                       // the partner's half is in its exchange buffer
                auto oth = [&](auto kc) {
                    constexpr int K = decltype(kc)::value;
                    cplx o[4], go[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        o[rr] = xother[(K * 4 + rr) * 32 + lane];
                        go[rr] = g[((K * 4 + rr) * 4 + g_oth) * 32];
                    }
                    mac_oth<H * 16 + K * 4>(X, o, go);
                };
                oth(std::integral_constant<int, 0>{});
                oth(std::integral_constant<int, 1>{});
                oth(std::integral_constant<int, 2>{});
                oth(std::integral_constant<int, 3>{});
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                if (++stage == NH) { stage = 0; phase ^= 1; }
                // the partner must be done with this warp's exchange buffer before it is overwritten: by the second
                // half's exchange store (H = 0), or by the transpose store after the next pass (H = 1, barrier there)
                if (H == 0) pair_barrier(1 + ctl);
            };
            half(std::integral_constant<int, 0>{});
            half(std::integral_constant<int, 1>{});
        };
        auto do_tail = [&]() {
            stream_tail<AccT>(lane, acc, tabs + kTabTwist, X);
            __syncwarp();
        };
        // one copy of the pass in the loop body (the body has to fit the instruction cache): q selects the table and
        // the stages around the pass
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            if (q == 0) do_head();
            else if (q & 1) xp_in(q == 1 ? lane : row_inv);
            // (pass32_uniform for q = 0, 2 inside this rolled loop was measured: 5.10 -> 6.02 ms per 296 blocks - three pass bodies
            // behind a switch, 255 registers and spills; the specialisation pays in the straight-line pbs_stream_tx_kernel<2> only)
            pass32(X, pass_table(tabs, q, lane));
            if (!(q & 1)) xp_out(q == 2);
            else if (q == 1) do_mac();
            else do_tail();
        }
        FSC_POLL();
    }
#undef FSC_POLL
    pair_barrier(1 + ctl);

    if (live) {
        const size_t out = (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
        const pair_t<AccT>* mask = acc_all + (size_t)(ctl * 2) * 1024;
        for (int j = (p * 32 + lane); j <= kN; j += 64) store_out_word(out_big, out + j, extract_word<AccT>(mask, mask + 1024, j));
    }
}

// ---------------------------------------------------------------------------------------
// Wide-batch variant with tensor memory (32-bit accumulator, 4 ciphertexts per CTA).  The two warps of a ciphertext
// are w and w + 4: they share a TMEM lane quarter, so everything that one thread hands to "the same lane" — of its
// partner warp or of itself later — goes through tcgen05.st / tcgen05.ld instead of the shared-memory pipe:
//   columns [0, 256)    spectra for the partner exchange (warp p writes [128 p, 128 p + 128), position order)
//   columns [256, 384)  own-index accumulator pairs of warp p at [256 + 64 p, ...), in the tail's position order
//   columns [384, 512)  the lane's 32 twist constants (written once by the p = 0 warp)
// Shared memory keeps what crosses lanes: the accumulator copy for the rotated reads, the transposes, the key ring,
// the per-lane pass tables.  One pair barrier hands the spectra over, one guards their reuse.
// ---------------------------------------------------------------------------------------
// MODE 0: one pass routine for the four passes (rolled trips).  MODE 1: the two uniform passes through pass32_uniform
// (constants from the constant bank, trivial constants as additions: 2 688 -> 2 532 FP64 instructions), trips rolled with a
// switch on the pass; MODE 2: the same, straight-line trips.
#ifndef FSC_TX_I2F
#define FSC_TX_I2F 1     // head of MODE 2: int -> double through the conversion unit (1) or the mantissa trick (0)
#endif
#ifndef FSC_TX_F2I
#define FSC_TX_F2I 1     // tail of MODE 2: rounding to the accumulator through the conversion unit (1) or to_torus32 (0)
#endif
#ifndef FSC_TX_ST
#define FSC_TX_ST 2      // spectrum stores of MODE 2: 2 = one double per tcgen05.st, 4 = one complex, 16 = four complex (gathered)
#endif
// AccT = uint64_t (MODE 2 only): the reference's accumulator width with the same four ciphertexts per SM.  The accumulator lives
// in tensor memory (128 columns per warp at [256, 512): where the 32-bit form keeps its accumulator and the twist constants), the
// by-index copy in the transpose buffer takes 16 KB of the 16.5 KB buffer and is complete whenever the head runs (the portable
// cmux_head reads rotated and own pairs from it; the transposes destroy it afterwards, the tail rewrites it), and the twist
// constants come from global memory (L2-resident table, requested one group ahead).
template <int MODE, typename AccT = uint32_t>
__global__ void __launch_bounds__(256, 1) pbs_stream_tx_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                                 int n, int base_log, const uint64_t* __restrict__ luts,
                                                                 const uint32_t* __restrict__ lut_idx, const __grid_constant__ OutDest out_big,
                                                                 const int32_t* __restrict__ out_idx, int count,
                                                                 const cplx* __restrict__ tabs_g, int stagger) {
    constexpr bool ACC64 = sizeof(AccT) == 8;
    static_assert(!ACC64 || MODE == 2, "the 64-bit accumulator exists in the straight-line form only");
    constexpr int CTS = 4, NH = 2;
    constexpr int kTmAcc = 256, kTmTwist = 384, kTmemCols = 512;
    constexpr int kTabNoTwist = kTabTwist;               // the twist table lives in tensor memory
    // MODE 2: the transposes move whole complex values through a [32][33] complex buffer per warp (one store phase, one load
    // phase, 16 bytes per lane and instruction, conflict-free through the padded row), and the by-index copy of the
    // accumulator polynomial for the rotated reads lives in the first 8 KB of that buffer (written by the tail, when the
    // buffer is free; private to the warp): no separate accumulator array, same footprint.
    constexpr bool WIDE = MODE == 2;
    constexpr int kXC = 32 * 33;                         // complex per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pair_t<AccT>* acc_all = reinterpret_cast<pair_t<AccT>*>(smem_raw);
    double* xbuf_all = reinterpret_cast<double*>(smem_raw + (WIDE ? 0 : (size_t)CTS * 2 * 1024 * sizeof(pair_t<AccT>)));
    cplx* ring = WIDE ? reinterpret_cast<cplx*>(smem_raw) + (size_t)CTS * 2 * kXC
                      : reinterpret_cast<cplx*>(xbuf_all + (size_t)CTS * 2 * kXBufDoubles);
    cplx* tabs = ring + (size_t)NH * kHalfCplx;
    uint64_t* full = reinterpret_cast<uint64_t*>(tabs + kTabNoTwist);
    uint64_t* empty = full + NH;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NH);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NH; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 2 * CTS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < kTabNoTwist; t += CTS * 64) {
        const double2 d = __ldg(reinterpret_cast<const double2*>(tabs_g + t));
        tabs[t].x = d.x; tabs[t].y = d.y;
        if (MODE == 2 && t >= kTabL1 && t < kTabL1 + 32) tabs[t].y = d.y / d.x;      // pass 1, level 1: (cos, tan) for pass32<true>
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_halves = 2 * n;
    const bool producer = warp == 0;                   // warp-uniform
    HalfProducer<NH> prod;
    prod.init();
    if (producer) prod.poll(lane, bsk_f, ring, full, empty, total_halves);

    const int ctl = warp & 3, p = warp >> 2;
    const int c_raw = blockIdx.x * CTS + ctl;
    const bool live = c_raw < count;
    const int c = live ? c_raw : count - 1;            // padding warps shadow the last ciphertext, never store
    cplx* xc = reinterpret_cast<cplx*>(smem_raw) + (size_t)(ctl * 2 + p) * kXC;      // MODE 2
    pair_t<AccT>* acc = WIDE ? reinterpret_cast<pair_t<AccT>*>(xc) : acc_all + (size_t)(ctl * 2 + p) * 1024;
    double* xb = xbuf_all + (size_t)(ctl * 2 + p) * kXBufDoubles;
    const uint64_t* ct = in_small + (size_t)c * (n + 1);
    const uint64_t* lut = luts + (size_t)(lut_idx ? lut_idx[c] : 0) * kN;
    const uint32_t t_quarter = tmem_base + ((uint32_t)(ctl * 32) << 16);
    const uint32_t t_own = t_quarter + (uint32_t)(p * 128), t_oth = t_quarter + (uint32_t)((1 - p) * 128);
    const uint32_t t_acc = t_quarter + kTmAcc + (uint32_t)(p * 64), t_tw = t_quarter + kTmTwist;
    const uint32_t t_acc64 = t_quarter + kTmAcc + (uint32_t)(p * 128);      // 64-bit accumulator: [position][x lo, x hi, y lo, y hi]
    (void)t_acc64;

    if (p == 0 && !ACC64) {      // the lane's twist constants, position order: 8 x 16 columns
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cplx v4[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const double2 d = __ldg(reinterpret_cast<const double2*>(tabs_g + kTabTwist + (k * 4 + rr) * 32 + lane));
                v4[rr].x = d.x; v4[rr].y = d.y;
            }
            tmem_st4(t_tw + 16 * k, v4);
        }
    }
    if constexpr (ACC64) {      // accumulator <- (0, X^{-b} LUT): by index in the warp's transpose buffer, position order in tensor memory
        const int b = modswitch(ct[n]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            uint32_t w[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = lane + 32 * tail_j2(4 * q + u);
                pair_t<AccT> z; z.x = 0; z.y = 0;
                if (p) z = lut_pair<AccT>(lut, idx, b);
                acc[idx] = z;
                w[4 * u] = (uint32_t)z.x; w[4 * u + 1] = (uint32_t)(z.x >> 32); w[4 * u + 2] = (uint32_t)z.y; w[4 * u + 3] = (uint32_t)(z.y >> 32);
            }
            tmem_stw16(t_acc64 + 16 * q, w);
        }
    } else {   // accumulator <- (0, X^{-b} LUT): shared memory by index, tensor memory in the tail's position order
        const int b = modswitch(ct[n]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t w[16];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = lane + 32 * tail_j2(8 * k + u);
                pair_t<AccT> z; z.x = 0; z.y = 0;
                if (p) z = lut_pair<AccT>(lut, idx, b);
                acc[idx] = z;
                w[2 * u] = z.x; w[2 * u + 1] = z.y;
            }
            tmem_stw16(t_acc + 16 * k, w);
        }
    }
    tmem_wait_st();
    tmem_fence_before();
    pair_barrier(1 + ctl);      // the twist constants staged by the p = 0 warp are visible to its partner
    tmem_fence_after();

    const int g_own = 3 * p, g_oth = 2 - p;
    const int row_inv = (32 - lane) & 31;
    int a_chunk = 0;
    int stage = 0;
    uint32_t phase = 0;
    cplx X[32];
    if (stagger != 0) {      // experiment (FSC_TX_STAGGER): ciphertexts of a CTA start at different times
        const long long t0 = clock64();
        // stagger > 0: ciphertext k waits k * stagger; stagger < 0: two groups (ciphertexts 0, 1 | 2, 3), the second -stagger late
        const long long wait = stagger > 0 ? (long long)stagger * ctl : (long long)(-stagger) * (ctl >> 1);
        while (clock64() - t0 < wait) { }
        __syncwarp();
    }
    for (int i = 0; i < n; ++i) {
        if ((i & 31) == 0) a_chunk = (i + lane < n) ? modswitch(ct[i + lane]) : 0;
        const int a = __shfl_sync(0xffffffffu, a_chunk, i & 31);

        auto do_head = [&]() {
            if constexpr (ACC64) { cmux_head<AccT>(lane, acc, a, base_log, X); return; }
            // stream_head_u32 with the own-index pairs from tensor memory; elements visited in the tail's position order
            const int sh = 32 - base_log;
            const int half = 1 << (sh - 1);
            const int base = (lane - a) & 4095;
            const int q0 = base >> 10, q1 = (q0 + 1) & 3;
            const int swA = q0 & 1;
            const int sxA = 1 - (q0 & 2), syA = 1 - ((q0 ^ (q0 << 1)) & 2);
            const int sxB = 1 - (q1 & 2), syB = 1 - ((q1 ^ (q1 << 1)) & 2);
            const int dsx = sxB - sxA, dsy = syB - syA;
            unsigned b8 = (unsigned)(base & 1023) << 3;
            const char* pb = reinterpret_cast<const char*>(acc);
            // MODE 2: the own-index pairs of chunk k + 1 are requested before chunk k is worked on (tcgen05.wait::ld waits for
            // every outstanding load, so the request goes out right after the wait: its latency hides behind 8 elements)
            uint32_t wbuf[2][16];
            if constexpr (MODE == 2) tmem_ldw16(t_acc, wbuf[0]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t (&w)[16] = wbuf[k & 1];
                if constexpr (MODE == 2) {
                    tmem_wait_ld();
                    if (k < 3) tmem_ldw16(t_acc + 16 * (k + 1), wbuf[(k + 1) & 1]);
                } else {
                    tmem_ldw16(t_acc + 16 * k, w);
                    tmem_wait_ld();
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j2 = tail_j2(8 * k + u);
                    if ((u & 3) == 0) asm volatile("" : "+r"(b8));
                    const unsigned uu = b8 + 256u * j2;
                    const int cc = (int)(uu >> 13);
                    const uint2 P = *reinterpret_cast<const uint2*>(pb + (uu & 8191u));
                    const int sw = swA ^ cc;
                    const int sx = imad(cc, dsx, sxA), sy = imad(cc, dsy, syA);
                    const int e = (int)(P.y - P.x);
                    const int px = imad(sw, e, (int)P.x);
                    const int py = (int)(P.x + P.y) - px;
                    const int dx = imad(px, sx, half - (int)w[2 * u]);
                    const int dy = imad(py, sy, half - (int)w[2 * u + 1]);
                    if constexpr (MODE == 2 && FSC_TX_I2F) {
                        // wide batches are bound by the FP64 pipe: the conversion unit (quarter rate, otherwise idle) takes the
                        // 64 int -> double conversions instead of 64 DADDs + 128 integer instructions of the mantissa trick
                        X[j2].x = (double)(dx >> sh);
                        X[j2].y = (double)(dy >> sh);
                    } else {
                        X[j2].x = __hiloint2double(0x43300000, (dx >> sh) ^ (int)0x80000000) - 4503601774854144.0;
                        X[j2].y = __hiloint2double(0x43300000, (dy >> sh) ^ (int)0x80000000) - 4503601774854144.0;
                    }
                }
            }
        };
        auto xp_out = [&]() {
            if constexpr (WIDE) {
                __syncwarp();      // the head's rotated reads of the by-index copy are done in every lane
#pragma unroll
                for (int pos = 0; pos < 32; ++pos) xc[brev5(pos) * 33 + lane] = X[pos];
            } else {
                xp_store(lane, xb, X, 0);
            }
            __syncwarp();
        };
        auto xp_in = [&](int row) {
            if constexpr (WIDE) {
#pragma unroll
                for (int s2 = 0; s2 < 32; ++s2) X[s2] = xc[row * 33 + s2];
                __syncwarp();
            } else {
                xp_load(row, xb, X, 0);
                __syncwarp();
                xp_store(lane, xb, X, 1);
                __syncwarp();
                xp_load(row, xb, X, 1);
                __syncwarp();
            }
        };
        auto do_mac = [&]() {
            if (producer) {      // both halves of this step requested before anyone sleeps on them (see pbs_stream_kernel)
                while (prod.next_h < 2 * (i + 1) && prod.next_h < total_halves)
                    prod.poll(lane, bsk_f, ring, full, empty, total_halves);
            }
            if constexpr (MODE != 2) {
                pair_barrier(1 + ctl);                       // the partner has read the spectrum of the previous step
                tmem_fence_after();
            }      // MODE 2: that barrier sits at the end of the previous product, where the two warps are in step anyway, so
                   // that the spectrum stores below can be scheduled into the last butterfly level of the pass before them
#pragma unroll
            for (int k = 0; k < 8; ++k) {                    // position order: the partner reads 4 consecutive positions per load
                if constexpr (MODE == 2 && FSC_TX_ST == 4) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) tmem_st_c1(t_own + 16 * k + 4 * rr, X[brev5(freq_at(4 * k + rr))]);
                } else if constexpr (MODE == 2 && FSC_TX_ST == 2) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        tmem_st_d1(t_own + 16 * k + 4 * rr, X[brev5(freq_at(4 * k + rr))].x);
                        tmem_st_d1(t_own + 16 * k + 4 * rr + 2, X[brev5(freq_at(4 * k + rr))].y);
                    }
                } else {
                    cplx v4[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) v4[rr] = X[brev5(freq_at(4 * k + rr))];
                    tmem_st4(t_own + 16 * k, v4);
                }
            }
            tmem_wait_st();
            tmem_fence_before();
            const int st0 = stage;
            mbar_wait(full + stage, phase);
            if (++stage == NH) { stage = 0; phase ^= 1; }
            const int st1 = stage;
            mbar_wait(full + stage, phase);
            if (++stage == NH) { stage = 0; phase ^= 1; }
            const cplx* g0 = ring + (size_t)st0 * kHalfCplx + lane;
            const cplx* g1 = ring + (size_t)st1 * kHalfCplx + lane;
            auto own = [&](auto kc) {
                constexpr int K = decltype(kc)::value;       // chunk of 4 positions; K < 4: first key half
                const cplx* g = (K < 4 ? g0 : g1);
                cplx gw[4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) gw[rr] = g[(((K & 3) * 4 + rr) * 4 + g_own) * 32];
                mac_own<K * 4>(X, gw);
            };
            own(std::integral_constant<int, 0>{}); own(std::integral_constant<int, 1>{});
            own(std::integral_constant<int, 2>{}); own(std::integral_constant<int, 3>{});
            own(std::integral_constant<int, 4>{}); own(std::integral_constant<int, 5>{});
            own(std::integral_constant<int, 6>{}); own(std::integral_constant<int, 7>{});
            pair_barrier(1 + ctl);                           // the partner's spectrum is in tensor memory
            tmem_fence_after();
            uint32_t obuf[2][16];
            if constexpr (MODE == 2) tmem_ldw16(t_oth, obuf[0]);
            auto oth = [&](auto kc) {
                constexpr int K = decltype(kc)::value;
                const cplx* g = (K < 4 ? g0 : g1);
                cplx o[4], go[4];
                if constexpr (MODE == 2) {
                    tmem_wait_ld();
                    if (K < 7) tmem_ldw16(t_oth + 16 * (K + 1), obuf[(K + 1) & 1]);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr)
                        o[rr] = cplx_from_words(obuf[K & 1][4 * rr], obuf[K & 1][4 * rr + 1], obuf[K & 1][4 * rr + 2], obuf[K & 1][4 * rr + 3]);
                } else {
                    tmem_ld4(t_oth + 16 * K, o);
                }
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) go[rr] = g[(((K & 3) * 4 + rr) * 4 + g_oth) * 32];
                mac_oth<K * 4>(X, o, go);
            };
            oth(std::integral_constant<int, 0>{}); oth(std::integral_constant<int, 1>{});
            oth(std::integral_constant<int, 2>{}); oth(std::integral_constant<int, 3>{});
            oth(std::integral_constant<int, 4>{}); oth(std::integral_constant<int, 5>{});
            oth(std::integral_constant<int, 6>{}); oth(std::integral_constant<int, 7>{});
            tmem_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
            if constexpr (MODE == 2) {
                pair_barrier(1 + ctl);                       // both warps have read each other's spectrum: it may be overwritten
                tmem_fence_after();
            }
        };
        auto do_tail = [&]() {
            // twist constants and own-index pairs from tensor memory; results to shared memory (rotated reads) and back
            if constexpr (ACC64) {
                // groups of 4 positions: accumulator words from tensor memory and twist constants from global memory one group ahead
                const double2* twg = reinterpret_cast<const double2*>(tabs_g + kTabTwist) + lane;
                uint32_t ab[2][16];
                double2 tg[2][4];
                tmem_ldw16(t_acc64, ab[0]);
#pragma unroll
                for (int u = 0; u < 4; ++u) tg[0][u] = __ldg(twg + u * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t (&w)[16] = ab[q & 1];
                    tmem_wait_ld();
                    if (q < 7) {
                        tmem_ldw16(t_acc64 + 16 * (q + 1), ab[(q + 1) & 1]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) tg[(q + 1) & 1][u] = __ldg(twg + (4 * (q + 1) + u) * 32);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double2 t = tg[q & 1][u];
                        const cplx x = X[4 * q + u];
                        pair_t<AccT> O;
                        O.x = ((uint64_t)w[4 * u + 1] << 32 | w[4 * u]) + to_acc_scaled<AccT>(fma(-x.y, t.y, x.x * t.x));
                        O.y = ((uint64_t)w[4 * u + 3] << 32 | w[4 * u + 2]) + to_acc_scaled<AccT>(fma(x.y, t.x, x.x * t.y));
                        acc[lane + 32 * tail_j2(4 * q + u)] = O;
                        w[4 * u] = (uint32_t)O.x; w[4 * u + 1] = (uint32_t)(O.x >> 32); w[4 * u + 2] = (uint32_t)O.y; w[4 * u + 3] = (uint32_t)(O.y >> 32);
                    }
                    tmem_stw16(t_acc64 + 16 * q, w);
                }
                tmem_wait_st();
                __syncwarp();
                return;
            } else if constexpr (MODE == 2) {
                // groups of 4 positions, the loads of group q + 1 in flight while group q is twisted, rounded and accumulated
                uint32_t tb[2][16], ab[2][8];
                tmem_ldw16(t_tw, tb[0]);
                tmem_ldw8(t_acc, ab[0]);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t (&tw)[16] = tb[q & 1];
                    uint32_t (&w)[8] = ab[q & 1];
                    tmem_wait_ld();
                    if (q < 7) { tmem_ldw16(t_tw + 16 * (q + 1), tb[(q + 1) & 1]); tmem_ldw8(t_acc + 8 * (q + 1), ab[(q + 1) & 1]); }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const cplx t = cplx_from_words(tw[4 * u], tw[4 * u + 1], tw[4 * u + 2], tw[4 * u + 3]);
                        const cplx x = X[4 * q + u];
                        const double re = fma(-x.y, t.y, x.x * t.x);
                        const double im = fma(x.y, t.x, x.x * t.y);
                        pair_t<AccT> O;
                        // round(v) mod 2^32 through the conversion unit: one F2I.S64 (|v| is 2^57 rms, 2^63 is 60 sigma away)
                        // instead of the four DADDs of to_torus32 - the same integer, and the FP64 pipe is what bounds this kernel
                        O.x = w[2 * u] + (FSC_TX_F2I ? (uint32_t)(uint64_t)__double2ll_rn(re) : to_acc_scaled<AccT>(re));
                        O.y = w[2 * u + 1] + (FSC_TX_F2I ? (uint32_t)(uint64_t)__double2ll_rn(im) : to_acc_scaled<AccT>(im));
                        acc[lane + 32 * tail_j2(4 * q + u)] = O;
                        w[2 * u] = O.x; w[2 * u + 1] = O.y;
                    }
                    tmem_stw8(t_acc + 8 * q, w);
                }
                tmem_wait_st();
                __syncwarp();
                return;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t tw[32], w[16];
                tmem_ldw32(t_tw + 32 * k, tw);
                tmem_ldw16(t_acc + 16 * k, w);
                tmem_wait_ld();
                double re[8], im[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const cplx t = cplx_from_words(tw[4 * u], tw[4 * u + 1], tw[4 * u + 2], tw[4 * u + 3]);
                    const cplx x = X[8 * k + u];
                    re[u] = fma(-x.y, t.y, x.x * t.x);
                    im[u] = fma(x.y, t.x, x.x * t.y);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    pair_t<AccT> O;
                    O.x = w[2 * u] + to_acc_scaled<AccT>(re[u]);
                    O.y = w[2 * u + 1] + to_acc_scaled<AccT>(im[u]);
                    acc[lane + 32 * tail_j2(8 * k + u)] = O;
                    w[2 * u] = O.x; w[2 * u + 1] = O.y;
                }
                tmem_stw16(t_acc + 16 * k, w);
            }
            tmem_wait_st();
            __syncwarp();
        };
        if constexpr (MODE == 2) {
            do_head();
            pass32_uniform<32>(X, UniConsts<0>{});
            xp_out();
            xp_in(lane);
            pass32<true>(X, pass_table(tabs, 1, lane));
            do_mac();
            pass32_uniform<0>(X, UniConsts<1>{});
            xp_out();
            xp_in(row_inv);
            pass32(X, pass_table(tabs, 3, lane));
            do_tail();
        } else {
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                if (q == 0) do_head();
                else if (q & 1) xp_in(q == 1 ? lane : row_inv);
                if (MODE == 1 && q == 0) pass32_uniform<32>(X, UniConsts<0>{});
                else if (MODE == 1 && q == 2) pass32_uniform<0>(X, UniConsts<1>{});
                else pass32(X, pass_table(tabs, q, lane));
                if (!(q & 1)) xp_out();
                else if (q == 1) do_mac();
                else do_tail();
            }
        }
        if (producer) prod.poll(lane, bsk_f, ring, full, empty, total_halves);
    }
    pair_barrier(1 + ctl);

    if (live) {
        const size_t out = (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
        const pair_t<AccT>* mask = WIDE ? reinterpret_cast<const pair_t<AccT>*>(reinterpret_cast<cplx*>(smem_raw) + (size_t)(ctl * 2) * kXC)
                                        : acc_all + (size_t)(ctl * 2) * 1024;
        const pair_t<AccT>* body = WIDE ? reinterpret_cast<const pair_t<AccT>*>(reinterpret_cast<cplx*>(smem_raw) + (size_t)(ctl * 2 + 1) * kXC)
                                        : mask + 1024;
        for (int j = (p * 32 + lane); j <= kN; j += 64) store_out_word(out_big, out + j, extract_word<AccT>(mask, body, j));
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) { tmem_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

// ---------------------------------------------------------------------------------------
// Test hook: c = a (torus) * b (small integers), negacyclic, through the stream FFT.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) negacyclic_mul_stream_kernel(const uint64_t* __restrict__ a, const int64_t* __restrict__ b,
                                                                   uint64_t* __restrict__ c, const cplx* __restrict__ tabs) {
    __shared__ double xb[kXBufDoubles];
    __shared__ pair_t<uint64_t> acc[1024];
    const int lane = threadIdx.x;
    a += (size_t)blockIdx.x * kN; b += (size_t)blockIdx.x * kN; c += (size_t)blockIdx.x * kN;
    cplx va[32], vb[32];
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        va[j2].x = (double)(int64_t)a[lane + 32 * j2]; va[j2].y = (double)(int64_t)a[lane + 32 * j2 + 1024];
        vb[j2].x = (double)b[lane + 32 * j2];          vb[j2].y = (double)b[lane + 32 * j2 + 1024];
        acc[lane + 32 * j2].x = 0; acc[lane + 32 * j2].y = 0;
    }
    stream_fft_fwd(lane, xb, tabs, va);
    __syncwarp();
    stream_fft_fwd(lane, xb, tabs, vb);
    __syncwarp();
    cplx pr[32];
#pragma unroll
    for (int s = 0; s < 32; ++s) {      // slot s holds k1 = brev5(s): the product goes to slot k1
        const cplx x = va[s], y = vb[s];
        pr[brev5(s)].x = x.x * y.x - x.y * y.y; pr[brev5(s)].y = x.x * y.y + x.y * y.x;
    }
    pass32(pr, pass_table(tabs, 2, lane));
    xp_store(lane, xb, pr, 0);
    __syncwarp();
    xp_load((32 - lane) & 31, xb, pr, 0);
    __syncwarp();
    xp_store(lane, xb, pr, 1);
    __syncwarp();
    xp_load((32 - lane) & 31, xb, pr, 1);
    __syncwarp();
    pass32(pr, pass_table(tabs, 3, lane));
    stream_tail<uint64_t>(lane, acc, tabs + kTabTwist, pr);
    __syncwarp();
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) { c[lane + 32 * j2] = acc[lane + 32 * j2].x; c[lane + 32 * j2 + 1024] = acc[lane + 32 * j2].y; }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
void launch_bsk_convert_stream(const uint64_t* bsk, void* out, int n, cudaStream_t st) {
    bsk_convert_stream_kernel<<<n * 4, 32, 0, st>>>(bsk, reinterpret_cast<cplx*>(out), stream_tables<uint64_t>());
}

void launch_negacyclic_mul_stream(const uint64_t* a, const int64_t* b, uint64_t* c, int count, cudaStream_t st) {
    negacyclic_mul_stream_kernel<<<count, 32, 0, st>>>(a, b, c, stream_tables<uint64_t>());
}

template <typename AccT, int CTS, int NH>
static void launch_pbs_stream_t(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                                const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    const size_t smem = (size_t)CTS * 2 * 1024 * sizeof(pair_t<AccT>) + (size_t)CTS * 2 * kXBufDoubles * sizeof(double) +
                        (size_t)NH * kHalfCplx * sizeof(cplx) + (size_t)kTabCplx * sizeof(cplx) + 2 * NH * sizeof(uint64_t);
    ensure_dynamic_smem(reinterpret_cast<const void*>(&pbs_stream_kernel<AccT, CTS, NH>), smem);      // per device (the opt-in is a per-device attribute)
    const int grid = (count + CTS - 1) / CTS;
    static const int stagger = [] { const char* e = getenv("FSC_STREAM_STAGGER"); return e ? atoi(e) : 0; }();
    pbs_stream_kernel<AccT, CTS, NH><<<grid, CTS * 64, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log, luts,
                                                                          lut_idx, out_big, out_idx, count, stream_tables<AccT>(), stagger);
}

template <typename AccT>
static void launch_pbs_stream_tx(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                                 const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    static const int mode = [] { const char* e = getenv("FSC_STREAM_TX"); return e ? atoi(e) : 2; }();      // 0 / 1: comparison forms
#define FSC_TX(...) do { \
    const size_t smem = ((mode == 2 || sizeof(AccT) == 8) ? (size_t)4 * 2 * 32 * 33 * sizeof(cplx) \
                                   : (size_t)4 * 2 * 1024 * sizeof(pair_t<uint32_t>) + (size_t)4 * 2 * kXBufDoubles * sizeof(double)) + \
                        (size_t)2 * kHalfCplx * sizeof(cplx) + (size_t)kTabTwist * sizeof(cplx) + 2 * 2 * sizeof(uint64_t) + 16; \
    ensure_dynamic_smem(reinterpret_cast<const void*>(&pbs_stream_tx_kernel<__VA_ARGS__>), smem); \
    pbs_stream_tx_kernel<__VA_ARGS__><<<(count + 3) / 4, 256, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log, luts, lut_idx, \
                                                                   out_big, out_idx, count, stream_tables<AccT>(), stagger); } while (0)
    static const int stagger = [] { const char* e = getenv("FSC_TX_STAGGER"); return e ? atoi(e) : 0; }();
    if (mode || sizeof(AccT) == 8) stream_uniform_constants_init();
    if constexpr (sizeof(AccT) == 8) { FSC_TX(2, uint64_t); }
    else { if (mode == 2) FSC_TX(2); else if (mode == 1) FSC_TX(1); else FSC_TX(0); }
#undef FSC_TX
}

// Configuration by batch width: wide levels pack 4 (u32 accumulator) or 3 (u64) ciphertexts per CTA so that one key
// chunk feeds them all; levels of at most one or two ciphertexts per SM use 1 or 2 per CTA (the latency-bound case of
// the carry-propagation levels: a less contended SM per ciphertext).
void launch_pbs_stream(int acc_bits, const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                       const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, int sm_count, cudaStream_t st) {
    if (count <= 0) return;
#define FSC_STREAM(ACC, CTS, NH) \
    launch_pbs_stream_t<ACC, CTS, NH>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st)
    if (acc_bits == 32) {
        if (count <= sm_count) FSC_STREAM(uint32_t, 1, 3);
        else if (count <= 2 * sm_count) FSC_STREAM(uint32_t, 2, 3);
        else if (getenv("FSC_STREAM_NO_TMEM")) FSC_STREAM(uint32_t, 4, 2);
        else launch_pbs_stream_tx<uint32_t>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st);
    } else {
        // 64-bit accumulator: two ciphertexts per CTA is what fits beside a whole-step ring (the product default for
        // 64-bit accumulators is the ring kernel of pbs_kernel.cu, three per CTA)
        if (count <= sm_count) FSC_STREAM(uint64_t, 1, 3);
        else if (count <= 2 * sm_count || getenv("FSC_STREAM_NO_TMEM")) FSC_STREAM(uint64_t, 2, 2);
        else launch_pbs_stream_tx<uint64_t>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st);
    }
#undef FSC_STREAM
}

}  // namespace fsc
