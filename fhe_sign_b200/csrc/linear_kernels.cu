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

// buffer[b] = pool[src[b]]  (collecting the blocks of a radix value for one contiguous download)
__global__ void __launch_bounds__(256) gather_kernel(const uint64_t* __restrict__ pool, const int32_t* __restrict__ src_idx,
                                                      uint64_t* __restrict__ buffer, int words) {
    const int b = blockIdx.x;
    const uint64_t* src = pool + (size_t)src_idx[b] * words;
    uint64_t* o = buffer + (size_t)b * words;
    for (int w = threadIdx.x; w < words; w += blockDim.x) o[w] = src[w];
}
void launch_gather(const uint64_t* pool, const int32_t* src_idx, uint64_t* buffer, int count, int words, cudaStream_t st) {
    if (count <= 0) return;
    gather_kernel<<<count, 256, 0, st>>>(pool, src_idx, buffer, words);
}

// ---------------------------------------------------------------------------------------
// Flag barrier between the ranks of a node over peer-mapped memory (level sharding, radix_cuda.cu).
// Lane q tells peer q "rank `rank` has reached barrier `seq`" - every kernel this rank launched before, including the
// blind rotation whose epilogue stored into the peers' pools, has completed (stream order), and the release at system
// scope publishes those stores - then waits until peer q's arrival shows up in the local flags.  The spin is bounded:
// on timeout the barrier records `seq` in *err and returns (reported as FSC_ERR_COMM at the next download).
// ---------------------------------------------------------------------------------------
struct PeerFlags {
    uint64_t* remote[kMaxPeers];      // remote[q] = flag array inside rank q's pool allocation (remote[rank] is local)
    uint64_t* err;
    int rank, world;
};
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerFlags pf, uint64_t seq, long long timeout_cycles) {
    const int q = threadIdx.x;
    if (q >= pf.world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pf.remote[q] + pf.rank), "l"(seq) : "memory");
    const uint64_t* mine = pf.remote[pf.rank] + q;
    const long long t0 = clock64();
    for (;;) {
        uint64_t v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= seq) break;
        if (clock64() - t0 > timeout_cycles) { *pf.err = seq; break; }
        __nanosleep(200);
    }
}
void launch_peer_barrier(uint64_t* const* remote_flags, uint64_t* err, int rank, int world, uint64_t seq, cudaStream_t st) {
    PeerFlags pf;
    for (int i = 0; i < kMaxPeers; ++i) pf.remote[i] = i < world ? remote_flags[i] : nullptr;
    pf.err = err; pf.rank = rank; pf.world = world;
    peer_barrier_kernel<<<1, 32, 0, st>>>(pf, seq, 8000000000ll);      // about 4 s at the B200's 1.9 GHz
}

void launch_lincomb(const uint64_t* pool, const int32_t* row_ptr, const int32_t* slot, const int32_t* coef, const int32_t* cst,
                    uint64_t delta, uint64_t* out, const int32_t* dst_idx, int count, int words, cudaStream_t st) {
    if (count <= 0) return;
    lincomb_kernel<<<count, 256, 0, st>>>(pool, row_ptr, slot, coef, cst, delta, out, dst_idx, words);
}

}  // namespace fsc
