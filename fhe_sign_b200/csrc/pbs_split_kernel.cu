// pbs_split_kernel.cu — programmable bootstrapping for sm_100a (k = 1, N = 2048, l = 1), latency form for narrow levels.
//
// A PBS level of at most one ciphertext per SM (the carry-propagation levels of every radix operator: 16 to 128 blocks)
// is n sequential CMUX steps on an otherwise idle SM.  pbs_stream_kernel gives such a ciphertext two warps, one per SM
// sub-partition, and half of the SM's FP64 pipes idle; this kernel gives it four — warp (p, h) owns slots
// [16 h, 16 h + 16) of GLWE polynomial p in every pass (pbs_core3.cuh) — so every sub-partition issues half of the
// instructions per step.  Same tables, same Fourier key layout and ring as the stream kernel; everything that crosses
// warps goes through shared memory and five block barriers per step:
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

namespace fsc {

template <typename AccT, int NH>
__global__ void __launch_bounds__(128, 1) pbs_split_kernel(const cplx* __restrict__ bsk_f, const uint64_t* __restrict__ in_small,
                                                            int n, int base_log, const uint64_t* __restrict__ luts,
                                                            const uint32_t* __restrict__ lut_idx, uint64_t* __restrict__ out_big,
                                                            const int32_t* __restrict__ out_idx, int count,
                                                            const cplx* __restrict__ tabs_g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pair_t<AccT>* acc_all = reinterpret_cast<pair_t<AccT>*>(smem_raw);
    cplx* E_all = reinterpret_cast<cplx*>(smem_raw + (size_t)2 * 1024 * sizeof(pair_t<AccT>));
    cplx* T_all = E_all + 2 * kSplitECplx;
    cplx* ring = T_all + 2 * kSplitTCplx;
    cplx* tabs = ring + (size_t)NH * kHalfCplx;
    uint64_t* full = reinterpret_cast<uint64_t*>(tabs + kTabCplx);
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

        split_head<AccT>(lane, h, acc, a, base_log, E);
        __syncthreads();
        FSC_POLL();
        split_pass(h, SplitLoadE{E + lane}, c0, w);
        split_xp_store(lane, h, T, w);
        __syncthreads();
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
            split_pass(h, ld, c2, w);
            __syncwarp();
            if (lane == 0) { mbar_arrive(empty + st0); mbar_arrive(empty + st1); }
        }
        split_xp_store(lane, h, T, w);
        __syncthreads();
        FSC_POLL();
        split_pass(h, SplitLoadT{T + row_inv * kSplitTRow}, c3, w);
        split_tail<AccT>(lane, h, acc, tabs + kTabTwist, w);
        __syncthreads();
    }
#undef FSC_POLL

    uint64_t* out = out_big + (size_t)(out_idx ? out_idx[c] : c) * (kN + 1);
    for (int j = threadIdx.x; j <= kN; j += 128) out[j] = extract_word<AccT>(acc_all, acc_all + 1024, j);
}

template <typename AccT, int NH>
static void launch_pbs_split_t(const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                               const uint32_t* lut_idx, uint64_t* out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    const size_t smem = (size_t)2 * 1024 * sizeof(pair_t<AccT>) + (size_t)2 * (kSplitECplx + kSplitTCplx) * sizeof(cplx) +
                        (size_t)NH * kHalfCplx * sizeof(cplx) + (size_t)kTabCplx * sizeof(cplx) + 2 * NH * sizeof(uint64_t);
    static bool configured = false;
    if (!configured) {
        FSC_CUDA_CHECK(cudaFuncSetAttribute(pbs_split_kernel<AccT, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    pbs_split_kernel<AccT, NH><<<count, 128, smem, st>>>(reinterpret_cast<const cplx*>(bsk_f), in_small, n, base_log, luts, lut_idx,
                                                         out_big, out_idx, count, stream_tables<AccT>());
}

// bsk_f: the stream kernel's Fourier key (launch_bsk_convert_stream).  One CTA per ciphertext: meant for count <= SMs.
void launch_pbs_split(int acc_bits, const void* bsk_f, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                      const uint32_t* lut_idx, uint64_t* out_big, const int32_t* out_idx, int count, cudaStream_t st) {
    if (count <= 0) return;
    if (acc_bits == 32) launch_pbs_split_t<uint32_t, 3>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st);
    else launch_pbs_split_t<uint64_t, 2>(bsk_f, in_small, n, base_log, luts, lut_idx, out_big, out_idx, count, st);
}

}  // namespace fsc
