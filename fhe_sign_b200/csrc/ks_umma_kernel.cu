// ks_umma_kernel.cu — the keyswitch limb GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Same exact contraction as ks_mma_kernel.cu,
//     S[c][8 col + b] = sum_k digit_k(c) * limb_b(KSK[k][col])      (s8 x u8 -> s32, cannot overflow)
//     out[c][col]     = (0,..,0,b_c)[col] - sum_b 2^(8b) S[c][8 col + b]                      (mod 2^64)
// with the operands already in the layouts the tensor cores want: digits D [C_pad][K] and limbs B [8(n+1) padded][K],
// both K-contiguous ("K-major").  One CTA computes a 128 x 256 tile of S:
//   warp 0     producer: TMA 2-D tile loads (128 B swizzle) of a 128 x 128 B slab of D and a 256 x 128 B slab of B per stage
//   warp 1     one elected lane issues tcgen05.mma.kind::i8 (M 128, N 256, K 32) four times per stage, accumulator in
//              TMEM (128 lanes x 256 columns of s32); tcgen05.commit releases the stage / publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld 32 columns at a time (thread = accumulator row), 8 limb columns -> one u64 word,
//              negate, add the body, store
// The mma.sync kernel (ks_mma_kernel.cu) stays as FSC_KS_VARIANT=mma and as a bit-exact cross-check.
//
// Replaces (concept): tfhe 0.10.0 keyswitch_lwe_ciphertext (Cargo.lock:482-485).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "fsc_internal.h"

namespace fsc {

constexpr int UM_BM = 128, UM_BN = 256, UM_BK = 128, UM_STAGES = 4, UM_THREADS = 192;
constexpr uint32_t UM_A_BYTES = UM_BM * UM_BK, UM_B_BYTES = UM_BN * UM_BK;
constexpr uint32_t UM_TMEM_COLS = 256;

__device__ __forceinline__ uint32_t um_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void um_mbar_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(um_smem(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void um_mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(um_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void um_mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "UM_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra UM_DONE;\n\t"
        "bra UM_WAIT;\n\t"
        "UM_DONE:\n\t"
        "}" ::"r"(um_smem(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void um_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(um_smem(dst)), "l"(map), "r"(c0), "r"(c1), "r"(um_smem(bar)) : "memory");
}
// K-major operand tile with 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), descriptor version 1
__device__ __forceinline__ uint64_t um_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // start address, 16-byte units
    d |= (uint64_t)1 << 16;                                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
    return d;
}
// kind::i8: D s32, A signed 8-bit, B unsigned 8-bit, both K-major, N >> 3, M >> 4
constexpr uint32_t UM_IDESC = (2u << 4) | (1u << 7) | (0u << 10) | ((UM_BN >> 3) << 17) | ((UM_BM >> 4) << 24);

__device__ __forceinline__ void um_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(UM_IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void um_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem(bar)) : "memory");
}
__device__ __forceinline__ bool um_elect() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

__global__ void __launch_bounds__(UM_THREADS, 1) ks_umma_kernel(const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_b,
                                                                 const uint64_t* __restrict__ in_big, uint64_t* __restrict__ out_small,
                                                                 int count, int K, int big_dim, int n) {
    extern __shared__ __align__(1024) unsigned char um_smem_raw[];
    // 1024-byte alignment of the swizzled tiles
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(um_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;                                        // [STAGES][128 x 128 B]
    unsigned char* sB = base + UM_STAGES * UM_A_BYTES;               // [STAGES][256 x 128 B]
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + UM_STAGES * UM_B_BYTES);
    uint64_t* empty = full + UM_STAGES;
    uint64_t* acc_ready = empty + UM_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * UM_BN;
    const int KT = K / UM_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < UM_STAGES; ++s) { um_mbar_init(full + s, 1); um_mbar_init(empty + s, 1); }
        um_mbar_init(acc_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {      // TMEM allocation by one full warp; the address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem(tmem_slot)), "n"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ---- producer ----
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % UM_STAGES;
            if (kt >= UM_STAGES) um_mbar_wait(empty + s, (uint32_t)((kt / UM_STAGES) - 1) & 1);
            if (um_elect()) {
                um_mbar_expect(full + s, UM_A_BYTES + UM_B_BYTES);
                um_tma_2d(sA + (size_t)s * UM_A_BYTES, &map_d, kt * UM_BK, m0, full + s);
                um_tma_2d(sB + (size_t)s * UM_B_BYTES, &map_b, kt * UM_BK, n0, full + s);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % UM_STAGES;
            um_mbar_wait(full + s, (uint32_t)(kt / UM_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (um_elect()) {
                const uint64_t da = um_smem_desc(um_smem(sA + (size_t)s * UM_A_BYTES));
                const uint64_t db = um_smem_desc(um_smem(sB + (size_t)s * UM_B_BYTES));
#pragma unroll
                for (int k = 0; k < UM_BK / 32; ++k)      // K = 32 bytes per instruction: +2 in 16-byte address units
                    um_mma_i8(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), (kt | k) != 0);
                um_commit(empty + s);                                   // stage reusable once these MMAs have read it
                if (kt == KT - 1) um_commit(acc_ready);                 // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue: warps 2..5 own TMEM lane quarters (warp % 4) ----
        um_mbar_wait(acc_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                 // accumulator row = TMEM lane
        const int c = m0 + row;
        const uint64_t body = (c < count) ? in_big[(size_t)c * (big_dim + 1) + big_dim] : 0;
#pragma unroll 1
        for (int cb = 0; cb < UM_BN / 32; ++cb) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int col = (n0 + cb * 32) / 8 + w;      // output word
                uint64_t sum = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) sum += (uint64_t)(int64_t)(int32_t)v[w * 8 + b] << (8 * b);
                if (c < count && col <= n) {
                    uint64_t o = (uint64_t)0 - sum;
                    if (col == n) o += body;
                    out_small[(size_t)c * (n + 1) + col] = o;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(UM_TMEM_COLS) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*um_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static um_encode_fn um_encoder() {
    static um_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FSC_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) throw Error(FSC_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<um_encode_fn>(p);
    }
    return fn;
}

// byte matrix [rows][K], K contiguous; box = 128 bytes x box_rows, 128-byte swizzle
static CUtensorMap um_make_map(const void* ptr, uint64_t rows, uint64_t K, uint32_t box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {K, rows};
    const cuuint64_t strides[1] = {K};
    const cuuint32_t box[2] = {(cuuint32_t)UM_BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = um_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(FSC_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

size_t ks_umma_limb_rows(int n) { return (size_t)((8 * (n + 1) + UM_BN - 1) / UM_BN) * UM_BN; }

// digits: [rows_pad][K] s8 written by ks_decompose_kernel (launched by the caller); limbs: [limb_rows][K] u8
void launch_keyswitch_umma(const uint8_t* limbs, size_t limb_rows, const int8_t* digits, int rows_pad, const uint64_t* in_big,
                           uint64_t* out_small, int count, int big_dim, int n, int level, cudaStream_t st) {
    if (count <= 0) return;
    const int K = big_dim * level;
    FSC_REQUIRE(K % UM_BK == 0 && rows_pad % UM_BM == 0 && limb_rows % UM_BN == 0, "keyswitch (tcgen05 path): unsupported shape");
    const CUtensorMap md = um_make_map(digits, (uint64_t)rows_pad, (uint64_t)K, UM_BM);
    const CUtensorMap mb = um_make_map(limbs, (uint64_t)limb_rows, (uint64_t)K, UM_BN);
    const size_t smem = (size_t)UM_STAGES * (UM_A_BYTES + UM_B_BYTES) + 1024 + 256;
    ensure_dynamic_smem(reinterpret_cast<const void*>(&ks_umma_kernel), smem);      // per device (the opt-in is a per-device attribute)
    dim3 grid((unsigned)(limb_rows / UM_BN), (unsigned)(rows_pad / UM_BM));
    ks_umma_kernel<<<grid, UM_THREADS, smem, st>>>(md, mb, in_big, out_small, count, K, big_dim, n);
}

}  // namespace fsc
