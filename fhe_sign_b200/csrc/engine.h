// engine.h — the per-context device state behind fsc_ctx (one CUDA stream, keys, scratch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "fsc_internal.h"

namespace fsc {

struct Luts {
    size_t n = 0;
    uint64_t* d = nullptr;      // [n][N] accumulator polynomials
};

void build_lut_poly(const fsc_params& p, const uint64_t* table, uint64_t* poly);

struct Engine {
    fsc_params p;
    int dev = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* bsk_f = nullptr;          // Fourier bootstrapping key [n][32 r][4 g][32 lane] complex f64, r in the order of pbs_variant
    void* bsk_s = nullptr;          // second copy in the stream kernel's order (pbs_variant 3 only)
    int pbs_variant = 3;            // 0 pair, 1 ring, 2 stream, 3 ring for wide batches + stream / split for narrow levels, 4 split: decided at key upload
    bool wide_tx = true;            // variant 3, 32-bit accumulator: wide batches run on pbs_stream_tx_kernel<2> (FSC_PBS_WIDE=ring: pbs_ring_kernel)
    bool use_split = true;          // variant 3: levels of at most one ciphertext per SM run on pbs_split_kernel (FSC_PBS_SPLIT=0: stream kernel)
    uint64_t* ksk = nullptr;        // [kN][l_ks][n+1]
    uint8_t* ksk_limbs = nullptr;   // byte-limb transpose of the KSK for the tensor-core keyswitch [8(n+1) padded][kN*l_ks]
    int8_t* ks_digits = nullptr;    // digit matrix of the current batch [rows padded to 128][kN*l_ks]
    size_t ks_digits_cap = 0;
    int ks_variant = 2;             // 0 = CUDA-core kernel, 1 = mma.sync s8 x u8 limb GEMM, 2 = tcgen05 kind::i8 limb GEMM (default)
    uint64_t* scratch_small = nullptr;   // keyswitch outputs of the current batch
    uint32_t* scratch_idx = nullptr;     // LUT indices of the current batch
    size_t scratch_cap = 0;
    uint64_t* scratch_big = nullptr;     // host-buffer entry point staging (in | out)
    size_t scratch_big_cap = 0;
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    uint64_t launches = 0;
    // host-buffer entry point: two copy streams so that uploads and downloads overlap the bootstraps of other chunks
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaStream_t lane[2] = {nullptr, nullptr};      // blind rotations of consecutive chunks of fsc_apply_lut_host: the next chunk's CTAs start as SMs free up
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::vector<cudaEvent_t> chunk_events;
    bool host_chunk_explicit = false;    // FSC_HOST_CHUNK_WAVES given: uniform chunks instead of the first-wave | middle | last-wave plan
    size_t host_chunk_waves = 1;         // chunk = this many waves of the blind-rotation kernel (FSC_HOST_CHUNK_WAVES; 0 = no overlap)

    Engine(const fsc_params& prm, int device, uintptr_t ext_stream);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    void use();
    void upload_keys(const uint64_t* bsk_std, size_t bsk_words, const uint64_t* ksk_h, size_t ksk_words);
    void ensure_scratch(size_t count);
    void ensure_pinned(size_t bytes);
    void ensure_copy_streams();
    cudaEvent_t chunk_event(size_t i);
    const uint32_t* stage_lut_idx(const uint32_t* lut_idx, size_t count, const Luts* luts);
    void keyswitch(const uint64_t* in_big, uint64_t* out_small, size_t count);
    // dests: optional multi-destination output (the same block pool on every GPU of the node, fsc_peer_*); overrides out_big
    void pbs(const uint64_t* in_small, const Luts* luts, const uint32_t* lut_idx_dev, uint64_t* out_big, size_t count,
             const int32_t* out_idx_dev = nullptr, const OutDest* dests = nullptr);
    void ks_pbs(const uint64_t* in_big, const Luts* luts, const uint32_t* lut_idx_dev, uint64_t* out_big, size_t count);
};

}  // namespace fsc
