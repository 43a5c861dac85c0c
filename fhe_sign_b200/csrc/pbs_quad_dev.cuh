// pbs_quad_dev.cuh — device-only helpers shared by pbs_quad_kernel.cu and pbs_duo_kernel.cu: named barriers with the
// tensor-memory fences around them, complex <-> tensor-memory column moves, and the head of warp (p, h) in the ALU / FMA-pipe
// form (the portable statement is quad_head<AccT> in pbs_core4.cuh, which the CPU emulator runs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbs_core4.cuh"
#include "tma_ring.cuh"
#include "pbs_head.cuh"

namespace fsc {

constexpr int kQuadCts = 4;
constexpr int kQTmAcc = 0, kQTmSpec = 128, kQTmJoin = 384, kQTmCols = 512;

__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// publish tensor-memory stores to the sibling warps / see theirs
__device__ __forceinline__ void tmem_barrier(int id, int threads) {
    tmem_wait_st();
    tmem_fence_before();
    named_barrier(id, threads);
    tmem_fence_after();
}
// four complex doubles at v[0..4) <-> 16 consecutive columns
__device__ __forceinline__ void tmem_st_c4(uint32_t taddr, const cplx* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr),
                 "r"(__double2loint(v[0].x)), "r"(__double2hiint(v[0].x)), "r"(__double2loint(v[0].y)), "r"(__double2hiint(v[0].y)),
                 "r"(__double2loint(v[1].x)), "r"(__double2hiint(v[1].x)), "r"(__double2loint(v[1].y)), "r"(__double2hiint(v[1].y)),
                 "r"(__double2loint(v[2].x)), "r"(__double2hiint(v[2].x)), "r"(__double2loint(v[2].y)), "r"(__double2hiint(v[2].y)),
                 "r"(__double2loint(v[3].x)), "r"(__double2hiint(v[3].x)), "r"(__double2loint(v[3].y)), "r"(__double2hiint(v[3].y)) : "memory");
}
__device__ __forceinline__ void tmem_ld_c4_issue(uint32_t taddr, uint32_t (&w)[16]) { tmem_ldw16(taddr, w); }
__device__ __forceinline__ void words_to_c4(const uint32_t (&w)[16], cplx* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = cplx_from_words(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}
__device__ __forceinline__ void tmem_stw32(uint32_t taddr, const uint32_t (&w)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
                 "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]),
                 "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]),
                 "r"(w[29]), "r"(w[30]), "r"(w[31]) : "memory");
}

// head of warp (p, h) in the ALU / FMA-pipe form of stream_head_u32 (pbs_head.cuh): digits of X^a acc - acc at
// j2 = 16 b + 8 h + u -> v[8 b + u].  Rotated pairs from the by-index copy in shared memory, own pairs from tensor memory
// (quad_acc_col: for each b the pairs of even u and of odd u are 8 consecutive columns each).
__device__ __forceinline__ void quad_head_u32(int lane, int h, const pair_t<uint32_t>* scratch, uint32_t t_acc, int a, int base_log,
                                              cplx (&v)[16]) {
    const int sh = 32 - base_log;
    const int half = 1 << (sh - 1);
    const int base = (lane - a) & 4095;
    const int q0 = base >> 10, q1 = (q0 + 1) & 3;
    const int swA = q0 & 1;
    const int sxA = 1 - (q0 & 2), syA = 1 - ((q0 ^ (q0 << 1)) & 2);
    const int sxB = 1 - (q1 & 2), syB = 1 - ((q1 ^ (q1 << 1)) & 2);
    const int dsx = sxB - sxA, dsy = syB - syA;
    const unsigned b8 = ((unsigned)(base & 1023) << 3) + 2048u * h;      // + 256 j2 per element, bit 13 = crossed a multiple of 1024
    const char* pb = reinterpret_cast<const char*>(scratch);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        uint32_t Oe[8], Oo[8];
        tmem_ldw8(t_acc + 16 * b + 8 * h, Oe);               // u = 0, 2, 4, 6
        tmem_ldw8(t_acc + 32 + 16 * b + 8 * h, Oo);          // u = 1, 3, 5, 7
        uint2 Pv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) Pv[u] = *reinterpret_cast<const uint2*>(pb + ((b8 + 4096u * b + 256u * u) & 8191u));
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned uu = b8 + 4096u * b + 256u * u;
            const int c = (int)(uu >> 13);
            const uint2 P = Pv[u];
            const uint32_t Ox = (u & 1) ? Oo[2 * (u >> 1)] : Oe[2 * (u >> 1)];
            const uint32_t Oy = (u & 1) ? Oo[2 * (u >> 1) + 1] : Oe[2 * (u >> 1) + 1];
            const int sw = swA ^ c;
            const int sx = imad(c, dsx, sxA), sy = imad(c, dsy, syA);
            const int d = (int)(P.y - P.x);
            const int px = imad(sw, d, (int)P.x);
            const int py = (int)(P.x + P.y) - px;
            const int dx = imad(px, sx, half - (int)Ox);
            const int dy = imad(py, sy, half - (int)Oy);
            v[8 * b + u].x = (double)(dx >> sh);      // conversion unit (I2F.F64), not the FP64 pipe
            v[8 * b + u].y = (double)(dy >> sh);
        }
    }
}

}  // namespace fsc
