// pbs_solo_kernel.cu — wide-batch blind rotation, ONE WARP PER CIPHERTEXT (32-bit accumulator, k = 1, N = 2048, l = 1).
//
// The ring / stream kernels give a ciphertext two warps (one per GLWE polynomial) that sit on the same SM
// sub-partition and meet twice per CMUX step to swap spectra: the two warps of a sub-partition are therefore in lock
// step, and whenever they are both in a phase that does not use the FP64 pipe (rotated difference, transposes, exchange,
// rounding: 41 % of the step, profiles/README.md) the pipe idles.  Here a warp owns a WHOLE ciphertext and walks its two
// polynomials one after the other, so that
//   * no warp ever waits for a partner: the only coupling between the 8 ciphertexts of a CTA is the key ring, and that
//     has four half-step stages (two whole steps of slack);
//   * the two warps of a sub-partition belong to different ciphertexts and are started half a step apart (warps 4-7
//     begin `stagger` cycles late), so one is in its FP64 passes while the other is in its integer / transpose phases;
//   * the Fourier-domain product needs no exchange: the spectrum of polynomial 0 waits in tensor memory while the warp
//     transforms polynomial 1, then both products are formed chunk by chunk (output 1 stays in registers, output 0
//     replaces the spectrum in tensor memory and is read back after the first inverse transform).
// Memory:
//   tensor memory, per warp 256 columns of its lane quarter (warps w and w + 4 share a quarter):
//       [0, 64) accumulator polynomial 0 (own-index pairs, tail position order) | [64, 128) polynomial 1 | [128, 256) spectrum
//   shared memory: per warp one 8.25 KB transpose buffer, which doubles as the scratch for the rotated accumulator reads
//       (the head dumps the polynomial's pairs into it, then gathers them rotated: the accumulator has NO permanent
//       shared-memory copy, which is what makes room for 8 ciphertexts and a 128 KB key ring), the key ring
//       (4 x 32 KB bulk copies, tma_ring.cuh), the pass / twist tables of the stream formulation (pbs_core2.cuh).
// One step = four trips through one code path (polynomial 0 forward | polynomial 1 forward + product | polynomial 1
// inverse + accumulate | polynomial 0 inverse + accumulate), each trip two calls of the single 32-point pass routine.
// Key layout: the stream kernel's (launch_bsk_convert_stream).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>
#include <vector>
#include "pbs_core2.cuh"
#include "pbs_head.cuh"
#include "fsc_internal.h"
#include "tma_ring.cuh"
#include "pbs_stream_tables.cuh"

namespace fsc {

constexpr int kSoloCts = 8;      // ciphertexts = warps per CTA

// FP64 token of a sub-partition's two warps (w and w + 4), built from two named barriers: a warp waits for its turn before
// an FP64-dense phase and hands the turn over after it.  Without it the two warps phase-lock: whenever they share the FP64
// pipe the one behind catches up (it gets the whole pipe once the leader leaves for its transposes), so any offset
// collapses to lock step and the pipe idles during both warps' integer / shared-memory phases (measured with
// FSC_SOLO_TRACE: offsets of 10-100 cycles whatever the start stagger).  Strict alternation pins the offset at one phase.
__device__ __forceinline__ void turn_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void turn_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

template <int NH, int PP>
__global__ void __launch_bounds__(kSoloCts * 32, 1) pbs_solo_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                                    int n, int base_log, const uint64_t* __restrict__ luts,
                                                                    const uint32_t* __restrict__ lut_idx, const __grid_constant__ OutDest out_big,
                                                                    const int32_t* __restrict__ out_idx, int count,
                                                                    const cplx* __restrict__ tabs_g, int stagger, int pin, long long* __restrict__ trace) {
    typedef uint32_t AccT;
    constexpr int kTmSpec = 128, kTmemCols = 512;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* xbuf_all = reinterpret_cast<double*>(smem_raw);
    cplx* ring = reinterpret_cast<cplx*>(xbuf_all + (size_t)kSoloCts * kXBufDoubles);
    cplx* tabs = ring + (size_t)NH * kHalfCplx;
    uint64_t* full = reinterpret_cast<uint64_t*>(tabs + kTabCplx);
    uint64_t* empty = full + NH;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NH);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NH; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kSoloCts); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < kTabCplx; t += kSoloCts * 32) {
        const double2 d = __ldg(reinterpret_cast<const double2*>(tabs_g + t));
        tabs[t].x = d.x; tabs[t].y = d.y;
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_halves = 2 * n;
    const bool producer = warp == 0;                   // warp-uniform; warp 0 belongs to the group that starts first
    HalfProducer<NH> prod;
    prod.init();
    if (producer) prod.poll(lane, bsk_f, ring, full, empty, total_halves);

    const int c_raw = blockIdx.x * kSoloCts + warp;
    const bool live = c_raw < count;
    const int c = live ? c_raw : count - 1;            // padding warps shadow the last ciphertext, never store
    double* xb = xbuf_all + (size_t)warp * kXBufDoubles;
    pair_t<AccT>* pairs = reinterpret_cast<pair_t<AccT>*>(xb);      // rotation scratch: the 1024 pairs of one polynomial
    const uint64_t* ct = in_small + (size_t)c * (n + 1);
    const uint64_t* lut = luts + (size_t)(lut_idx ? lut_idx[c] : 0) * kN;
    const uint32_t t_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    const uint32_t t_spec = t_base + kTmSpec;

    {   // accumulator <- (0, X^{-b} LUT), own-index pairs in the tail's position order
        const int b = modswitch(ct[n]);
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t w[16];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = lane + 32 * tail_j2(8 * k + u);
                    pair_t<AccT> z; z.x = 0; z.y = 0;
                    if (p) z = lut_pair<AccT>(lut, idx, b);
                    w[2 * u] = z.x; w[2 * u + 1] = z.y;
                }
                tmem_stw16(t_base + 64 * p + 16 * k, w);
            }
        }
        tmem_wait_st();
    }
    // named barriers 1..8: (1 + 2 q) = turn of the first warp of sub-partition q, (2 + 2 q) = turn of the second
    const int my_turn = 1 + 2 * (warp & 3) + (warp >> 2), other_turn = 1 + 2 * (warp & 3) + (1 - (warp >> 2));
    if (PP && PP < 3 && warp >= 4) turn_pass(other_turn);      // the first warp of the sub-partition goes first
    // PP == 3: no turns; the two warps of a sub-partition are PINNED a fixed part of a step apart instead (one named barrier
    // per step: the first warp arrives before the first pass of trip 0, the second before pass pin & 1 of trip pin >> 1),
    // and share the FP64 pipe freely in between.
    const int pin_code = warp < 4 ? 0 : pin;
    const int pin_bar = 9 + (warp & 3);
    if (warp >= 4 && stagger > 0) {      // second warp of every sub-partition: half a step behind the first
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)stagger) { }
    }
    __syncwarp();

    const int row_inv = (32 - lane) & 31;
    const int sh = 32 - base_log;
    const int half_ulp = 1 << (sh - 1);
    int a_chunk = 0;
    int stage = 0;
    uint32_t phase = 0;
    cplx X[32];
    for (int i = 0; i < n; ++i) {
        if ((i & 31) == 0) a_chunk = (i + lane < n) ? modswitch(ct[i + lane]) : 0;
        const int a = __shfl_sync(0xffffffffu, a_chunk, i & 31);
        if (trace && blockIdx.x == 0 && lane == 0 && (i & 63) == 0) trace[(i >> 6) * kSoloCts + warp] = clock64();      // diagnostic (FSC_SOLO_TRACE)
        // diagnostic: phase boundaries of warps 0 and 4 at step 512 (32 stamps each, behind the per-step rows)
        long long* stamps = (trace && blockIdx.x == 0 && lane == 0 && i == 512 && (warp & 3) == 0) ? trace + (n / 64 + 1) * kSoloCts + (warp >> 2) * 32 : nullptr;
        int n_stamp = 0;
#define FSC_STAMP() do { if (stamps) stamps[n_stamp++] = clock64(); } while (0)
        FSC_STAMP();

        // rotation constants of this step (pbs_head.cuh stream_head_u32)
        const int base = (lane - a) & 4095;
        const int q0 = base >> 10, q1 = (q0 + 1) & 3;
        const int swA = q0 & 1;
        const int sxA = 1 - (q0 & 2), syA = 1 - ((q0 ^ (q0 << 1)) & 2);
        const int sxB = 1 - (q1 & 2), syB = 1 - ((q1 ^ (q1 << 1)) & 2);
        const int dsx = sxB - sxA, dsy = syB - syA;
        const unsigned b8_0 = (unsigned)(base & 1023) << 3;

#pragma unroll 1
        for (int t = 0; t < 4; ++t) {
            const int p = (t < 2) ? t : 3 - t;                      // polynomial of this trip: 0, 1, 1, 0
            const uint32_t t_acc = t_base + (uint32_t)(64 * p);
            if (producer) prod.poll(lane, bsk_f, ring, full, empty, total_halves);

            if (t < 2) {
                // ---- head: digits of X^a acc_p - acc_p.  The pairs go from tensor memory into the scratch, then come back rotated.
                // (Keeping the whole polynomial's 64 own words and all 32 rotated pairs in flight at once was measured: it
                // spills - 780 B against 400 B - and the step takes 45.5 k cycles instead of 34.9 k.)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t w[16];
                    tmem_ldw16(t_acc + 16 * k, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        pair_t<AccT> z; z.x = w[2 * u]; z.y = w[2 * u + 1];
                        pairs[lane + 32 * tail_j2(8 * k + u)] = z;
                    }
                }
                __syncwarp();
                unsigned b8 = b8_0;
                const char* pb = reinterpret_cast<const char*>(pairs);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t w[16];
                    tmem_ldw16(t_acc + 16 * k, w);
                    tmem_wait_ld();
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
                        const int dx = imad(px, sx, half_ulp - (int)w[2 * u]);
                        const int dy = imad(py, sy, half_ulp - (int)w[2 * u + 1]);
                        X[j2].x = __hiloint2double(0x43300000, (dx >> sh) ^ (int)0x80000000) - 4503601774854144.0;
                        X[j2].y = __hiloint2double(0x43300000, (dy >> sh) ^ (int)0x80000000) - 4503601774854144.0;
                    }
                }
                __syncwarp();      // the scratch is the transpose buffer of the passes below
                FSC_STAMP();       // head done
            } else if (t == 3) {
                // ---- output 0 of the product comes back from tensor memory (position order -> slot)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    cplx o4[4];
                    tmem_ld4(t_spec + 16 * k, o4);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) X[freq_at(4 * k + rr)] = o4[rr];
                }
                FSC_STAMP();       // spectrum reload done
            }

            // ---- two passes of the one routine with a transpose between them (forward: tables 0, 1; inverse: 2, 3)
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                if (PP == 3) { if (2 * t + h == pin_code) turn_wait(pin_bar); }
                else if (PP) turn_wait(my_turn);
                FSC_STAMP();       // turn taken
                pass32(X, pass_table(tabs, (t < 2 ? 0 : 2) + h, lane));
                if (PP && PP < 3) turn_pass(other_turn);
                FSC_STAMP();       // pass done
                if (h == 0) {
                    const int row = t < 2 ? lane : row_inv;
                    xp_store(lane, xb, X, 0);
                    __syncwarp();
                    xp_load(row, xb, X, 0);
                    __syncwarp();
                    xp_store(lane, xb, X, 1);
                    __syncwarp();
                    xp_load(row, xb, X, 1);
                    __syncwarp();
                    FSC_STAMP();   // transpose done
                }
            }

            if (t == 0) {
                // ---- spectrum of polynomial 0 waits in tensor memory (position order)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    cplx v4[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) v4[rr] = X[brev5(freq_at(4 * k + rr))];
                    tmem_st4(t_spec + 16 * k, v4);
                }
                tmem_wait_st();
                FSC_STAMP();       // spectrum parked
            } else if (t == 1) {
                // ---- Fourier-domain product: Y_q = X_0 G[0][q] + X_1 G[1][q]; Y_1 -> registers, Y_0 -> tensor memory
                if (producer) {      // both halves of this step requested before this warp sleeps on them
                    while (prod.next_h < 2 * (i + 1) && prod.next_h < total_halves)
                        prod.poll(lane, bsk_f, ring, full, empty, total_halves);
                }
                const int st0 = stage;
                mbar_wait(full + stage, phase);
                if (++stage == NH) { stage = 0; phase ^= 1; }
                const int st1 = stage;
                mbar_wait(full + stage, phase);
                if (++stage == NH) { stage = 0; phase ^= 1; }
                const cplx* g0 = ring + (size_t)st0 * kHalfCplx + lane;
                const cplx* g1 = ring + (size_t)st1 * kHalfCplx + lane;
                if (PP == 2) turn_wait(my_turn);
                auto chunk = [&](auto kc) {
                    constexpr int K = decltype(kc)::value;       // 4 positions of the consumption order; K < 4: first key half
                    const cplx* g = (K < 4 ? g0 : g1) + (size_t)((K & 3) * 4) * 4 * 32;
                    cplx x0[4], x1[4], y0[4];
                    tmem_ld4(t_spec + 16 * K, x0);
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) x1[rr] = X[brev5(freq_at(4 * K + rr))];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const cplx g00 = g[(rr * 4 + 0) * 32], g01 = g[(rr * 4 + 1) * 32];
                        const cplx g10 = g[(rr * 4 + 2) * 32], g11 = g[(rr * 4 + 3) * 32];
                        const cplx a0 = x0[rr], a1 = x1[rr];
                        y0[rr].x = fma(-a1.y, g10.y, fma(a1.x, g10.x, fma(-a0.y, g00.y, a0.x * g00.x)));
                        y0[rr].y = fma(a1.y, g10.x, fma(a1.x, g10.y, fma(a0.y, g00.x, a0.x * g00.y)));
                        cplx y1;
                        y1.x = fma(-a1.y, g11.y, fma(a1.x, g11.x, fma(-a0.y, g01.y, a0.x * g01.x)));
                        y1.y = fma(a1.y, g11.x, fma(a1.x, g11.y, fma(a0.y, g01.x, a0.x * g01.y)));
                        X[freq_at(4 * K + rr)] = y1;
                    }
                    tmem_st4(t_spec + 16 * K, y0);
                };
                chunk(std::integral_constant<int, 0>{}); chunk(std::integral_constant<int, 1>{});
                chunk(std::integral_constant<int, 2>{}); chunk(std::integral_constant<int, 3>{});
                chunk(std::integral_constant<int, 4>{}); chunk(std::integral_constant<int, 5>{});
                chunk(std::integral_constant<int, 6>{}); chunk(std::integral_constant<int, 7>{});
                if (PP == 2) turn_pass(other_turn);
                tmem_wait_st();
                __syncwarp();
                if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
                FSC_STAMP();       // product done
            } else {
                // ---- tail: twist, rounding, accumulation into the own-index pairs in tensor memory
                const cplx* tw = tabs + kTabTwist + lane;
                if (PP == 2) turn_wait(my_turn);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t w[16];
                    tmem_ldw16(t_acc + 16 * k, w);
                    double re[8], im[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const cplx tc = tw[(8 * k + u) * 32];
                        const cplx x = X[8 * k + u];
                        re[u] = fma(-x.y, tc.y, x.x * tc.x);
                        im[u] = fma(x.y, tc.x, x.x * tc.y);
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        w[2 * u] = w[2 * u] + to_acc_scaled<AccT>(re[u]);
                        w[2 * u + 1] = w[2 * u + 1] + to_acc_scaled<AccT>(im[u]);
                    }
                    tmem_stw16(t_acc + 16 * k, w);
                }
                if (PP == 2) turn_pass(other_turn);
                tmem_wait_st();
                FSC_STAMP();       // tail done
            }
        }
#undef FSC_STAMP
    }

    // ---- sample extraction of coefficient 0: the mask polynomial through the scratch, the body's coefficient 0 from lane 0
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t w[16];
            tmem_ldw16(t_base + 16 * k, w);
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                pair_t<AccT> z; z.x = w[2 * u]; z.y = w[2 * u + 1];
                pairs[lane + 32 * tail_j2(8 * k + u)] = z;
            }
        }
        uint32_t wb[16];
        tmem_ldw16(t_base + 64, wb);      // polynomial 1, positions 0..7: position 0 is pair index `lane` (tail_j2(0) = 0)
        tmem_wait_ld();
        __syncwarp();
        if (live) {
            const size_t out = (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
            for (int j = lane; j < kN; j += 32) store_out_word(out_big, out + j, extract_word<AccT>(pairs, pairs, j));
            if (lane == 0) store_out_word(out_big, out + kN, (uint64_t)wb[0] << 32);
        }
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) { tmem_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

template <int NH, int PP>
static void launch_pbs_solo_t(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts, const uint32_t* lut_idx,
                              const OutDest& out_big, const int32_t* out_idx, int count, int stagger, int pin, cudaStream_t st) {
    const size_t smem = (size_t)kSoloCts * kXBufDoubles * sizeof(double) + (size_t)NH * kHalfCplx * sizeof(cplx) +
                        (size_t)kTabCplx * sizeof(cplx) + 2 * NH * sizeof(uint64_t) + 16;
    ensure_dynamic_smem(reinterpret_cast<const void*>(&pbs_solo_kernel<NH, PP>), smem);
    const int grid = (count + kSoloCts - 1) / kSoloCts;
    static const bool want_trace = getenv("FSC_SOLO_TRACE") != nullptr;      // diagnostic: per-warp clock at every 64th step of CTA 0
    long long* trace = nullptr;
    const int rows = n / 64 + 1;
    const size_t trace_words = (size_t)rows * kSoloCts + 64;      // + 2 x 32 phase stamps
    if (want_trace) {
        FSC_CUDA_CHECK(cudaMalloc(&trace, trace_words * sizeof(long long)));
        FSC_CUDA_CHECK(cudaMemsetAsync(trace, 0, trace_words * sizeof(long long), st));
    }
    pbs_solo_kernel<NH, PP><<<grid, kSoloCts * 32, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log, luts, lut_idx,
                                                           out_big, out_idx, count, stream_tables<uint32_t>(), stagger, pin, trace);
    if (want_trace) {
        std::vector<long long> h(trace_words);
        FSC_CUDA_CHECK(cudaStreamSynchronize(st));
        FSC_CUDA_CHECK(cudaMemcpy(h.data(), trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(trace);
        for (int r = 0; r < rows; ++r) {
            fprintf(stderr, "solo trace step %4d:", r * 64);
            for (int w = 0; w < kSoloCts; ++w) fprintf(stderr, " %9lld", h[(size_t)r * kSoloCts + w] - h[(size_t)r * kSoloCts]);
            if (r) fprintf(stderr, "   (cycles per step of warp 0: %lld)", (h[(size_t)r * kSoloCts] - h[(size_t)(r - 1) * kSoloCts]) / 64);
            fprintf(stderr, "\n");
        }
        static const char* const kPhase[] = {"start", "head0", "turn", "pass", "xpose", "turn", "pass", "park", "head1", "turn", "pass", "xpose", "turn", "pass",
                                             "product", "turn", "pass", "xpose", "turn", "pass", "tail1", "reload", "turn", "pass", "xpose", "turn", "pass", "tail0"};
        for (int wq = 0; wq < 2; ++wq) {
            const long long* s0 = h.data() + (size_t)rows * kSoloCts + wq * 32;
            if (!s0[0]) continue;
            fprintf(stderr, "solo phases, warp %d, step 512 (cycles):", wq * 4);
            for (int k = 1; k < 28 && s0[k]; ++k) fprintf(stderr, " %s %lld", kPhase[k], s0[k] - s0[k - 1]);
            fprintf(stderr, "\n");
        }
    }
}

// bsk_f: the stream kernel's Fourier key.  32-bit accumulator only.  FSC_SOLO_STAGGER: start delay of warps 4-7 in cycles
// (default: half of a measured step); FSC_SOLO_NH: ring depth in half steps (3 or 4).
void launch_pbs_solo(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts, const uint32_t* lut_idx,
                     const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    if (count <= 0) return;
    static const int stagger = [] { const char* e = getenv("FSC_SOLO_STAGGER"); return e ? atoi(e) : 0; }();
    static const int nh = [] { const char* e = getenv("FSC_SOLO_NH"); return e ? atoi(e) : 4; }();
    static const int pp = [] { const char* e = getenv("FSC_SOLO_PP"); return e ? atoi(e) : 1; }();      // FP64 turn-taking: 0 off, 1 passes, 2 passes + product + tail
    static const int pin = [] { const char* e = getenv("FSC_SOLO_PIN"); return e ? atoi(e) : 4; }();      // PP == 3: 2 * trip + pass of the second warp's pin point
#define FSC_SOLO(NH, PP) launch_pbs_solo_t<NH, PP>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, stagger, pin, st)
    if (nh == 3) { if (pp == 0) FSC_SOLO(3, 0); else if (pp == 1) FSC_SOLO(3, 1); else if (pp == 2) FSC_SOLO(3, 2); else FSC_SOLO(3, 3); }
    else { if (pp == 0) FSC_SOLO(4, 0); else if (pp == 1) FSC_SOLO(4, 1); else if (pp == 2) FSC_SOLO(4, 2); else FSC_SOLO(4, 3); }
#undef FSC_SOLO
}

}  // namespace fsc
