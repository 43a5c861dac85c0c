// ks_kernel.cu — batched LWE keyswitch (big key, dimension kN = 2048 -> small key, dimension n).
//
//   out[c] = (0, ..., 0, b_c) - sum_i sum_l digit_{i,l}(a_{c,i}) * KSK[i][l]        (wrapping u64)
//
// Integer work, bit-exact by construction: the signed digits follow the same closest-representable /
// balanced-carry rule as the CPU oracle (orc_decompose under oracle/), and u64 wrapping sums are
// order independent.
//
// Tiling: a CTA owns TC ciphertexts x 128 output columns.  The KSK row element is read once per
// thread (coalesced 8 B per lane) and reused for the TC ciphertexts held in registers; the digits of
// a chunk of mask elements are produced cooperatively into shared memory as int8 and broadcast with
// one 16-byte load per (i, l).
//
// Replaces (concept): tfhe 0.10.0 keyswitch_lwe_ciphertext (Cargo.lock:482-485), the first half of
// shortint apply_lookup_table behind every operator in src/biguint.rs:110-248.
#include <cuda_runtime.h>
#include <stdint.h>
#include "fsc_internal.h"

namespace fsc {

constexpr int KS_TC = 16;        // ciphertexts per CTA
constexpr int KS_COLS = 128;     // output columns per CTA (= threads)
constexpr int KS_IC = 32;        // mask elements per shared-memory chunk
constexpr int KS_MAXL = 8;

// digits[l] multiplies q / B^(l+1); produced from the least significant level upwards
__device__ __forceinline__ void ks_decompose(uint64_t x, int base_log, int level, int8_t* digits, int stride) {
    const int rep = base_log * level;
    uint64_t state = ((x >> (64 - rep - 1)) + 1) >> 1;
    state &= (rep < 64) ? (((uint64_t)1 << rep) - 1) : ~(uint64_t)0;
    const uint64_t B = (uint64_t)1 << base_log;
    for (int l = level - 1; l >= 0; --l) {
        uint64_t d = state & (B - 1);
        state >>= base_log;
        const uint64_t carry = (((d - 1) | state) & d) >> (base_log - 1);
        state += carry;
        digits[l * stride] = (int8_t)((int64_t)d - (int64_t)(carry << base_log));
    }
}

__global__ void __launch_bounds__(KS_COLS) keyswitch_kernel(const uint64_t* __restrict__ ksk, const uint64_t* __restrict__ in_big,
                                                             uint64_t* __restrict__ out_small, int count, int big_dim, int n,
                                                             int base_log, int level) {
    __shared__ __align__(16) int8_t dig[KS_IC * KS_MAXL * KS_TC];
    const int col = blockIdx.x * KS_COLS + threadIdx.x;
    const int c0 = blockIdx.y * KS_TC;
    const int row = n + 1;
    const bool active = col < row;
    uint64_t acc[KS_TC];
#pragma unroll
    for (int t = 0; t < KS_TC; ++t) acc[t] = 0;

    for (int i0 = 0; i0 < big_dim; i0 += KS_IC) {
        __syncthreads();
        // cooperative decomposition of KS_IC mask elements of KS_TC ciphertexts
        for (int e = threadIdx.x; e < KS_IC * KS_TC; e += KS_COLS) {
            const int t = e / KS_IC, i = e % KS_IC;          // consecutive threads -> consecutive i (coalesced)
            const int c = c0 + t;
            const uint64_t a = (c < count && i0 + i < big_dim) ? in_big[(size_t)c * (big_dim + 1) + i0 + i] : 0;
            ks_decompose(a, base_log, level, dig + (size_t)(i * level) * KS_TC + t, KS_TC);
        }
        __syncthreads();
        if (active) {
            const uint64_t* kp = ksk + ((size_t)i0 * level) * row + col;
            const int lim = min(KS_IC, big_dim - i0) * level;
#pragma unroll 2
            for (int il = 0; il < lim; ++il) {
                const uint64_t k = __ldg(kp + (size_t)il * row);
                const int4 dv = *reinterpret_cast<const int4*>(dig + il * KS_TC);
                const int w[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int t = 0; t < KS_TC; ++t) {
                    const int d = (int)(int8_t)(w[t >> 2] >> (8 * (t & 3)));
                    acc[t] -= (uint64_t)((int64_t)d) * k;
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int t = 0; t < KS_TC; ++t) {
            const int c = c0 + t;
            if (c < count) {
                uint64_t v = acc[t];
                if (col == n) v += in_big[(size_t)c * (big_dim + 1) + big_dim];
                out_small[(size_t)c * row + col] = v;
            }
        }
    }
}

void launch_keyswitch(const uint64_t* ksk, const uint64_t* in_big, uint64_t* out_small, int count, int big_dim, int n,
                      int base_log, int level, cudaStream_t st) {
    if (count <= 0) return;
    FSC_REQUIRE(level <= KS_MAXL && base_log * level < 64 && base_log <= 7, "keyswitch: unsupported decomposition");
    dim3 grid((n + 1 + KS_COLS - 1) / KS_COLS, (count + KS_TC - 1) / KS_TC);
    keyswitch_kernel<<<grid, KS_COLS, 0, st>>>(ksk, in_big, out_small, count, big_dim, n, base_log, level);
}

}  // namespace fsc
