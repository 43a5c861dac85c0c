// fsc_internal.h — declarations shared by the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>

#include "../../include/fhe_sign_cuda.h"

namespace fsc {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FSC_CUDA_CHECK(expr)                                                                          \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            throw ::fsc::Error(e__ == cudaErrorMemoryAllocation ? FSC_ERR_OOM : FSC_ERR_CUDA,         \
                               std::string(#expr) + ": " + cudaGetErrorString(e__));                  \
    } while (0)

#define FSC_REQUIRE(cond, msg)                                          \
    do {                                                                \
        if (!(cond)) throw ::fsc::Error(FSC_ERR_BAD_ARG, (msg));        \
    } while (0)

// Where a blind rotation writes its output ciphertext: base[0] is the local destination array; with level sharding over
// peer-mapped block pools (radix_cuda.cu, fsc_peer_*) base[1 .. n) are the SAME pool on the other GPUs of the node, and the
// sample-extraction epilogue stores every word to all of them over NVLink - the exchange is fused into the kernel.
constexpr int kMaxPeers = 8;
struct OutDest {
    uint64_t* base[kMaxPeers];
    int n;
};
#ifdef __CUDACC__
// one extracted word to every destination pool (local store + NVLink peer stores; d.n is uniform across the grid)
__device__ __forceinline__ void store_out_word(const OutDest& d, size_t off, uint64_t w) {
    d.base[0][off] = w;
#pragma unroll 1
    for (int i = 1; i < d.n; ++i) d.base[i][off] = w;
}
#endif
inline OutDest single_dest(uint64_t* out_big) {
    OutDest d;
    for (int i = 0; i < kMaxPeers; ++i) d.base[i] = nullptr;
    d.base[0] = out_big; d.n = 1;
    return d;
}

// fsc_api.cu: opts `kernel` into `bytes` of dynamic shared memory on the CURRENT device, once per (kernel, device).
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: a process-wide flag would leave a second
// context on another GPU with the 48 KB default.  Thread-safe.
void ensure_dynamic_smem(const void* kernel, size_t bytes);

// pbs_kernel.cu
void pbs_init_constants();
void launch_bsk_convert(const uint64_t* bsk_std, void* bsk_fourier, int n, cudaStream_t st);
void launch_pbs(int variant, int acc_bits, const void* bsk_fourier, const uint64_t* in_small, int n, int base_log,
                const uint64_t* luts, const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count,
                int sm_count, cudaStream_t st);
void launch_negacyclic_mul(const uint64_t* a, const int64_t* b, uint64_t* c, int count, cudaStream_t st);

// bsk_exact.cu: the same two layouts, correctly rounded (direct DFT in double-double arithmetic); either output may be null
void launch_bsk_convert_exact(const uint64_t* bsk_std, void* out_ring, void* out_stream, int n, cudaStream_t st);

// pbs_stream_kernel.cu (single-routine formulation; its own Fourier key layout)
void launch_bsk_convert_stream(const uint64_t* bsk_std, void* bsk_fourier, int n, cudaStream_t st);
void launch_pbs_stream(int acc_bits, const void* bsk_fourier, const uint64_t* in_small, int n, int base_log,
                       const uint64_t* luts, const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count,
                       int sm_count, cudaStream_t st);
// pbs_split_kernel.cu (latency form for narrow levels: four warps per ciphertext; the stream kernel's key layout)
void launch_pbs_split(int acc_bits, const void* bsk_fourier, const uint64_t* in_small, int n, int base_log, const uint64_t* luts,
                      const uint32_t* lut_idx, const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st);
// pbs_solo_kernel.cu (wide batches, one warp per ciphertext, accumulator in tensor memory; the stream kernel's key layout; 32-bit accumulator)
void launch_pbs_solo(const void* bsk_fourier, const uint64_t* in_small, int n, int base_log, const uint64_t* luts, const uint32_t* lut_idx,
                     const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st);
void launch_negacyclic_mul_stream(const uint64_t* a, const int64_t* b, uint64_t* c, int count, cudaStream_t st);
// pbs_quad_kernel.cu (wide batches, four warps per ciphertext at 128 registers, sibling exchange through tensor memory; the stream kernel's key layout; 32-bit accumulator)
void launch_pbs_quad(const void* bsk_fourier, const uint64_t* in_small, int n, int base_log, const uint64_t* luts, const uint32_t* lut_idx,
                     const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st);
// pbs_duo_kernel.cu (wide batches, two warps per ciphertext, each carrying half of BOTH polynomials as two instruction streams one segment apart; 32-bit accumulator)
void launch_pbs_duo(const void* bsk_fourier, const uint64_t* in_small, int n, int base_log, const uint64_t* luts, const uint32_t* lut_idx,
                    const OutDest& out_big, const int32_t* out_idx, int count, cudaStream_t st);
// 7: duo (wide) + stream / split (narrow);  0: pair, 1: ring, 2: stream, 3: ring (wide) + stream / split (narrow), 4: split, 5: solo (wide) + stream / split (narrow), 6: quad (wide) + stream / split (narrow)
// (FSC_PBS_VARIANT, else by accumulator width);
// fixed per context at key upload
int pbs_variant_for(int acc_bits);

double launch_fp64_peak(double* sink, int sm_count, int iters, cudaStream_t st);

// ks_kernel.cu
void launch_keyswitch(const uint64_t* ksk, const uint64_t* in_big, uint64_t* out_small, int count, int big_dim,
                      int n, int base_log, int level, cudaStream_t st);

// ks_mma_kernel.cu (tensor-core keyswitch)
size_t ks_mma_limb_rows(int n);
size_t ks_mma_digit_rows(size_t count);
void launch_ksk_limb_transpose(const uint64_t* ksk, uint8_t* out, int K, int n, cudaStream_t st);
void launch_ks_decompose(const uint64_t* in_big, int8_t* digits, int count, int rows_pad, int big_dim, int base_log, int level,
                         cudaStream_t st);
// ks_umma_kernel.cu (tcgen05 + TMEM + TMA version of the same GEMM; shares the limb and digit matrices)
void launch_keyswitch_umma(const uint8_t* limbs, size_t limb_rows, const int8_t* digits, int rows_pad, const uint64_t* in_big,
                           uint64_t* out_small, int count, int big_dim, int n, int level, cudaStream_t st);
void launch_keyswitch_mma(const uint8_t* limbs, int8_t* digits, const uint64_t* in_big, uint64_t* out_small, int count,
                          int big_dim, int n, int base_log, int level, cudaStream_t st);

// linear_kernels.cu: out[dst(b)] = sum_t coef[t] * pool[slot[t]] + cst[b] * delta on the body
void launch_lincomb(const uint64_t* pool, const int32_t* row_ptr, const int32_t* slot, const int32_t* coef,
                    const int32_t* cst, uint64_t delta, uint64_t* out, const int32_t* dst_idx, int count, int words,
                    cudaStream_t st);

void launch_scatter(const uint64_t* buffer, const int32_t* dst_idx, uint64_t* pool, int count, int words, cudaStream_t st);
// flag barrier over peer-mapped memory: remote_flags[q] = flag array in rank q's pool allocation (q = rank: local)
void launch_peer_barrier(uint64_t* const* remote_flags, uint64_t* err, int rank, int world, uint64_t seq, cudaStream_t st);
void launch_gather(const uint64_t* pool, const int32_t* src_idx, uint64_t* buffer, int count, int words, cudaStream_t st);

}  // namespace fsc
