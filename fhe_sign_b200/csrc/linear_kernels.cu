// linear_kernels.cu — blockwise LWE linear combinations (the K4 kernels of SURVEY.md section 2):
// additions, subtractions, scalar multiplications, plaintext additions and the lhs*4 + rhs packing of
// bivariate lookups, for a whole PBS level in one launch.  Pure u64 wrapping arithmetic, HBM-bound:
// one CTA per output ciphertext, coalesced 8-byte accesses over the 2049 words.
//
// Replaces (concept): tfhe 0.10.0 lwe_ciphertext_add / cleartext_mul / plaintext_add (Cargo.lock:482-485).
#include <cuda_runtime.h>
#include <stdint.h>
#include "fsc_internal.h"

namespace fsc {

__global__ void __launch_bounds__(256) lincomb_kernel(const uint64_t* __restrict__ pool, const int32_t* __restrict__ row_ptr,
                                                       const int32_t* __restrict__ slot, const int32_t* __restrict__ coef,
                                                       const int32_t* __restrict__ cst, uint64_t delta, uint64_t* __restrict__ out,
                                                       const int32_t* __restrict__ dst_idx, int words) {
    const int b = blockIdx.x;
    const int t0 = row_ptr[b], t1 = row_ptr[b + 1];
    uint64_t* o = out + (size_t)(dst_idx ? dst_idx[b] : b) * words;
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
        uint64_t acc = (w == words - 1) ? (uint64_t)(int64_t)cst[b] * delta : 0;
        for (int t = t0; t < t1; ++t) acc += (uint64_t)(int64_t)coef[t] * pool[(size_t)slot[t] * words + w];
        o[w] = acc;
    }
}

// pool[dst[b]] = buffer[b]  (placing an all-gathered level into its pool slots)
__global__ void __launch_bounds__(256) scatter_kernel(const uint64_t* __restrict__ buffer, const int32_t* __restrict__ dst_idx,
                                                       uint64_t* __restrict__ pool, int words) {
    const int b = blockIdx.x;
    const uint64_t* src = buffer + (size_t)b * words;
    uint64_t* o = pool + (size_t)dst_idx[b] * words;
    for (int w = threadIdx.x; w < words; w += blockDim.x) o[w] = src[w];
}
void launch_scatter(const uint64_t* buffer, const int32_t* dst_idx, uint64_t* pool, int count, int words, cudaStream_t st) {
    if (count <= 0) return;
    scatter_kernel<<<count, 256, 0, st>>>(buffer, dst_idx, pool, words);
}

void launch_lincomb(const uint64_t* pool, const int32_t* row_ptr, const int32_t* slot, const int32_t* coef, const int32_t* cst,
                    uint64_t delta, uint64_t* out, const int32_t* dst_idx, int count, int words, cudaStream_t st) {
    if (count <= 0) return;
    lincomb_kernel<<<count, 256, 0, st>>>(pool, row_ptr, slot, coef, cst, delta, out, dst_idx, words);
}

}  // namespace fsc
