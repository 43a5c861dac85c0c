// ks_mma_kernel.cu — batched LWE keyswitch as an exact int8 tensor-core contraction.
//
//   out[c][col] = (0,..,0,b_c)[col] - sum_{k=(i,l)} digit_k(c) * KSK[k][col]          (mod 2^64)
//
// The signed digits are tiny (|d| <= 2^(base_log-1) <= 64) and every u64 key word is 8 unsigned byte
// limbs, so the sum is  sum_b 2^(8b) * ( sum_k d_k * limb_b(KSK[k][col]) )  with each inner sum an
// s8 x u8 -> s32 dot product that cannot overflow (10240 * 64 * 255 < 2^31).  That inner sum is a dense
// [C x K] x [K x 8(n+1)] integer GEMM, run here on the tensor cores with mma.sync.m16n8k32.s8.u8
// (B200 keeps the INT8 tensor path); the epilogue recombines the 8 limb columns of each output word
// with shifts inside the quad that holds them.  Integer arithmetic throughout: bit-exact.
//
//   ks_decompose_kernel   big LWE masks -> digit matrix D [C_pad][K] (s8, K contiguous)
//   ksk_limb_transpose    KSK [K][n+1] u64 -> limb matrix B [8(n+1) padded][K] (u8, K contiguous), at upload
//   ks_mma_kernel         128 x 128 x 64 tiles, 4-stage cp.async pipeline, swizzled ldmatrix
//
// Replaces (concept): tfhe 0.10.0 keyswitch_lwe_ciphertext (Cargo.lock:482-485).
#include <cuda_runtime.h>
#include <stdint.h>
#include "fsc_internal.h"

namespace fsc {

constexpr int MM_BM = 128, MM_BN = 128, MM_BK = 64, MM_STAGES = 4, MM_THREADS = 256;

// ---- digit matrix --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ks_decompose_kernel(const uint64_t* __restrict__ in_big, int8_t* __restrict__ D, int count,
                                                            int rows_pad, int big_dim, int base_log, int level) {
    const int K = big_dim * level;
    const size_t total = (size_t)rows_pad * big_dim;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(e / big_dim), i = (int)(e % big_dim);
        const uint64_t x = c < count ? in_big[(size_t)c * (big_dim + 1) + i] : 0;
        const int rep = base_log * level;
        uint64_t state = ((x >> (64 - rep - 1)) + 1) >> 1;
        state &= ((uint64_t)1 << rep) - 1;
        const uint64_t B = (uint64_t)1 << base_log;
        int8_t* d = D + (size_t)c * K + (size_t)i * level;
        for (int l = level - 1; l >= 0; --l) {
            uint64_t dg = state & (B - 1);
            state >>= base_log;
            const uint64_t carry = (((dg - 1) | state) & dg) >> (base_log - 1);
            state += carry;
            d[l] = (int8_t)((int64_t)dg - (int64_t)(carry << base_log));
        }
    }
}

// ---- key preparation (once per key upload) ------------------------------------------------------
// out[(col*8 + b) * K + k] = byte b of ksk[k * row + col]; rows >= 8*row are zero
__global__ void __launch_bounds__(256) ksk_limb_transpose_kernel(const uint64_t* __restrict__ ksk, uint8_t* __restrict__ out, int K, int row,
                                                                  int n_pad) {
    __shared__ uint64_t tile[32][33];
    const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int k = k0 + r, col = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (k < K && col < row) ? ksk[(size_t)k * row + col] : 0;
    }
    __syncthreads();
    // thread (x = k offset, y) writes bytes for column c0 + cc, limb b
    for (int idx = threadIdx.y; idx < 32 * 8; idx += 8) {
        const int cc = idx >> 3, b = idx & 7;
        const int nl = (c0 + cc) * 8 + b, k = k0 + threadIdx.x;
        if (nl < n_pad && k < K) out[(size_t)nl * K + k] = (uint8_t)(tile[threadIdx.x][cc] >> (8 * b));
    }
}

// ---- GEMM -------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void mma_s8u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// tile rows are 64 bytes = 4 chunks of 16 B; the chunk index is XOR-swizzled with (row >> 1) & 3
__device__ __forceinline__ int swz(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

__global__ void __launch_bounds__(MM_THREADS) ks_mma_kernel(const int8_t* __restrict__ D, const uint8_t* __restrict__ Bm,
                                                             const uint64_t* __restrict__ in_big, uint64_t* __restrict__ out_small,
                                                             int count, int K, int big_dim, int n) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;                                   // [STAGES][BM * 64]
    unsigned char* sB = smem + MM_STAGES * MM_BM * MM_BK;       // [STAGES][BN * 64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;                    // 2 x 4 warps, warp tile 64 x 32
    const int m0 = blockIdx.y * MM_BM, n0 = blockIdx.x * MM_BN;
    const int8_t* gA = D + (size_t)m0 * K;
    const uint8_t* gB = Bm + (size_t)n0 * K;

    auto load_stage = [&](int stage, int kt) {
        const int kbase = kt * MM_BK;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int e = tid + it * MM_THREADS;                // 512 chunks per operand tile
            const int row = e >> 2, chunk = e & 3;
            cp_async16(sA + stage * (MM_BM * MM_BK) + swz(row, chunk), gA + (size_t)row * K + kbase + chunk * 16);
            cp_async16(sB + stage * (MM_BN * MM_BK) + swz(row, chunk), gB + (size_t)row * K + kbase + chunk * 16);
        }
    };

    int acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0;

    const int KT = K / MM_BK;
#pragma unroll
    for (int s = 0; s < MM_STAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<MM_STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + MM_STAGES - 1;
            if (nk < KT) load_stage(nk % MM_STAGES, nk);
            cp_async_commit();
        }
        const unsigned char* a_st = sA + (kt % MM_STAGES) * (MM_BM * MM_BK);
        const unsigned char* b_st = sB + (kt % MM_STAGES) * (MM_BN * MM_BK);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t af[4][4], bf[2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = wm * 64 + i * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(af[i], a_st + swz(row, ks * 2 + (lane >> 4)));
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int row = wn * 32 + j * 16 + (lane & 7) + (lane >> 4) * 8;
                ldmatrix_x4(bf[j], b_st + swz(row, ks * 2 + ((lane >> 3) & 1)));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_s8u8(acc[i][j], af[i], bf[j >> 1][(j & 1) * 2], bf[j >> 1][(j & 1) * 2 + 1]);
        }
    }
    cp_async_wait<0>();

    // epilogue: one n8 tile = the 8 byte limbs of one output word; quad lane q holds limbs 2q, 2q+1
    const int q = lane & 3, rq = lane >> 2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = (n0 + wn * 32 + j * 8) >> 3;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = m0 + wm * 64 + i * 16 + rq + h * 8;
                uint64_t v = ((uint64_t)(int64_t)acc[i][j][h * 2] << (16 * q)) + ((uint64_t)(int64_t)acc[i][j][h * 2 + 1] << (16 * q + 8));
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (q == 0 && c < count && col <= n) {
                    uint64_t o = (uint64_t)0 - v;
                    if (col == n) o += in_big[(size_t)c * (big_dim + 1) + big_dim];
                    out_small[(size_t)c * (n + 1) + col] = o;
                }
            }
        }
}

// ---- host side ----------------------------------------------------------------------------------------
// padded to 256 rows: the limb matrix is shared with the tcgen05 kernel (ks_umma_kernel.cu, 256-column tiles)
size_t ks_mma_limb_rows(int n) { return (size_t)((8 * (n + 1) + 255) / 256) * 256; }
size_t ks_mma_digit_rows(size_t count) { return (count + MM_BM - 1) / MM_BM * MM_BM; }

void launch_ksk_limb_transpose(const uint64_t* ksk, uint8_t* out, int K, int n, cudaStream_t st) {
    const int row = n + 1, n_pad = (int)ks_mma_limb_rows(n);
    dim3 grid((K + 31) / 32, (n_pad / 8 + 31) / 32), block(32, 8);
    ksk_limb_transpose_kernel<<<grid, block, 0, st>>>(ksk, out, K, row, n_pad);
}

void launch_ks_decompose(const uint64_t* in_big, int8_t* digits, int count, int rows_pad, int big_dim, int base_log, int level,
                         cudaStream_t st) {
    ks_decompose_kernel<<<592, 256, 0, st>>>(in_big, digits, count, rows_pad, big_dim, base_log, level);
}

void launch_keyswitch_mma(const uint8_t* limbs, int8_t* digits, const uint64_t* in_big, uint64_t* out_small, int count, int big_dim,
                          int n, int base_log, int level, cudaStream_t st) {
    if (count <= 0) return;
    const int K = big_dim * level;
    FSC_REQUIRE(K % MM_BK == 0 && base_log <= 7 && base_log * level < 64, "keyswitch (tensor-core path): unsupported decomposition");
    const int rows_pad = (int)ks_mma_digit_rows(count);
    ks_decompose_kernel<<<592, 256, 0, st>>>(in_big, digits, count, rows_pad, big_dim, base_log, level);
    const size_t smem = (size_t)MM_STAGES * (MM_BM + MM_BN) * MM_BK;
    ensure_dynamic_smem(reinterpret_cast<const void*>(&ks_mma_kernel), smem);      // per device (the opt-in is a per-device attribute)
    dim3 grid((unsigned)(ks_mma_limb_rows(n) / MM_BN), (unsigned)(rows_pad / MM_BM));
    ks_mma_kernel<<<grid, MM_THREADS, smem, st>>>(digits, limbs, in_big, out_small, count, K, big_dim, n);
}

}  // namespace fsc
