// tma_ring.cuh — shared-memory ring fed by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx): the PTX
// wrappers and the non-blocking producer used by the blind-rotation kernels to stream the Fourier bootstrapping key.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbs_core.cuh"

namespace fsc {

constexpr int kChunkSlots = 4;                         // frequency slots per ring chunk
constexpr int kChunkCplx = kChunkSlots * 4 * 32;       // 512 complex = 8 KiB
constexpr int kChunksPerStep = 32 / kChunkSlots;       // 8

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pair_barrier(int id) {      // the two warps of one ciphertext
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}
// same with immediate barrier ids, so that ptxas reserves CTS + 1 named barriers instead of all 16 (two CTAs per SM)
template <int CTS>
__device__ __forceinline__ void pair_barrier_imm(int ctl) {
    if (CTS == 1 || ctl == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else if (CTS == 2 || ctl == 1) asm volatile("bar.sync 2, 64;" ::: "memory");
    else if (CTS == 3 || ctl == 2) asm volatile("bar.sync 3, 64;" ::: "memory");
    else asm volatile("bar.sync 4, 64;" ::: "memory");
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// Producer state is kept by every lane of warp 0 and advanced with warp-uniform control flow (the barrier test reads
// one address, so all lanes agree); only the two issue instructions are predicated on lane 0.  A producer that
// diverges from its warp makes that warp - a consumer like the others - run whole phases twice.
template <int NCH>
struct RingProducer {
    int next_u, stage;
    uint32_t phase;            // parity of the `empty` phase that frees `stage`
    __device__ __forceinline__ void init() { next_u = 0; stage = 0; phase = 1; }      // phase 1: "previous" phase of a fresh barrier
    __device__ __forceinline__ void poll(int lane, const cplx* bsk_f, cplx* ring, uint64_t* full, uint64_t* empty, int total_chunks) {
        while (next_u < total_chunks) {
            if (!mbar_test(empty + stage, phase)) break;
            if (lane == 0) {
                mbar_arrive_expect_tx(full + stage, kChunkCplx * sizeof(cplx));
                bulk_load(ring + (size_t)stage * kChunkCplx, bsk_f + (size_t)next_u * kChunkCplx, kChunkCplx * sizeof(cplx), full + stage);
            }
            __syncwarp();
            ++next_u;
            if (++stage == NCH) { stage = 0; phase ^= 1; }
        }
    }
};

// Half-step ring: one 32 KB bulk copy brings the GGSW entries of 16 frequencies
// (one half of the consumption order), NH such stages form the ring.  Producer state lives in every lane of warp 0
// and advances with warp-uniform control flow (tma_ring.cuh explains why); it never blocks.
constexpr int kHalfCplx = 16 * 4 * 32;                  // 2048 complex = 32 KiB
template <int NH>
struct HalfProducer {
    int next_h, stage;
    uint32_t phase;
    __device__ __forceinline__ void init() { next_h = 0; stage = 0; phase = 1; }
    __device__ __forceinline__ void poll(int lane, const cplx* bsk_f, cplx* ring, uint64_t* full, uint64_t* empty, int total_halves) {
        while (next_h < total_halves) {
            if (!mbar_test(empty + stage, phase)) break;
            if (lane == 0) {
                mbar_arrive_expect_tx(full + stage, kHalfCplx * sizeof(cplx));
                bulk_load(ring + (size_t)stage * kHalfCplx, bsk_f + (size_t)next_h * kHalfCplx, kHalfCplx * sizeof(cplx), full + stage);
            }
            __syncwarp();
            ++next_h;
            if (++stage == NH) { stage = 0; phase ^= 1; }
        }
    }
};

// ---- tensor memory (TMEM) as a lane-aligned exchange buffer ----------------------------------------------------
// Warps w and w + 4 of a CTA address the same 32 TMEM lanes (w % 4); thread t of either warp reaches lane t of that
// quarter.  What one of them stores with tcgen05.st the other reads back with tcgen05.ld at the same (lane, column):
// a register-to-register hand-over between the two warps that does not touch the shared-memory pipe.
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {          // one full warp; COLS: power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {                // one full warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// four complex doubles (16 x 32 bit) of the calling thread's lane <-> 16 consecutive columns
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const cplx (&v)[4]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[4 * i + 0] = (uint32_t)__double2loint(v[i].x); w[4 * i + 1] = (uint32_t)__double2hiint(v[i].x);
        w[4 * i + 2] = (uint32_t)__double2loint(v[i].y); w[4 * i + 3] = (uint32_t)__double2hiint(v[i].y);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
                 "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
}
// one complex double <-> 4 consecutive columns, one double <-> 2: the operands are the value's own register pair(s), so ptxas
// need not gather sixteen words into consecutive registers first (16 IMAD.MOV per x16 store, profiles/r02q_*)
__device__ __forceinline__ void tmem_st_c1(uint32_t taddr, const cplx& v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__double2loint(v.x)), "r"(__double2hiint(v.x)),
                 "r"(__double2loint(v.y)), "r"(__double2hiint(v.y)) : "memory");
}
__device__ __forceinline__ void tmem_st_d1(uint32_t taddr, double v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ void tmem_ld_d1(uint32_t taddr, uint32_t (&w)[2]) {      // tmem_wait_ld() before use
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(taddr) : "memory");
}
// raw 32-bit columns of the calling thread's lane
__device__ __forceinline__ void tmem_ldw4(uint32_t taddr, uint32_t (&w)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ldw8(uint32_t taddr, uint32_t (&w)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ldw16(uint32_t taddr, uint32_t (&w)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
                   "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ldw32(uint32_t taddr, uint32_t (&w)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
                   "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]),
                   "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]),
                   "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_stw16(uint32_t taddr, const uint32_t (&w)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
                 "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
}
__device__ __forceinline__ void tmem_stw8(uint32_t taddr, const uint32_t (&w)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ cplx cplx_from_words(uint32_t xl, uint32_t xh, uint32_t yl, uint32_t yh) {
    cplx c;
    c.x = __hiloint2double((int)xh, (int)xl);
    c.y = __hiloint2double((int)yh, (int)yl);
    return c;
}

// Per-lane node constants of a 32-point pass kept in tensor memory (16 complex = 64 columns of the lane): pass32 /
// pass32_inv_gs call begin_level(L) once per butterfly level and get(ci) per node; the level's 1, 1, 2, 4 or 8 constants
// come in with one tcgen05.ld instead of that many shared-memory loads.
struct TmemLaneConsts {
    uint32_t taddr;
    mutable cplx c[8];
    mutable int ci0;
    __device__ __forceinline__ void begin_level(int L) const {
        ci0 = (L == 1) ? 0 : (1 << (L - 2));
        const int nconst = (L <= 2) ? 1 : (1 << (L - 2));
        if (nconst == 1) {
            uint32_t w[4];
            tmem_ldw4(taddr + 4 * ci0, w);
            tmem_wait_ld();
            c[0] = cplx_from_words(w[0], w[1], w[2], w[3]);
        } else if (nconst == 2) {
            uint32_t w[8];
            tmem_ldw8(taddr + 4 * ci0, w);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 2; ++i) c[i] = cplx_from_words(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h * 4 < nconst) {
                    uint32_t w[16];
                    tmem_ldw16(taddr + 4 * ci0 + 16 * h, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[h * 4 + i] = cplx_from_words(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
                }
            }
        }
    }
    __device__ __forceinline__ cplx get(int ci) const { return c[ci - ci0]; }
};

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, cplx (&v)[4]) {      // tmem_wait_ld() before the values are used
    uint32_t w[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
                   "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "r"(taddr) : "memory");
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i].x = __hiloint2double((int)w[4 * i + 1], (int)w[4 * i + 0]);
        v[i].y = __hiloint2double((int)w[4 * i + 3], (int)w[4 * i + 2]);
    }
}

}  // namespace fsc
