// pbs_head.cuh — device-only head of a CMUX step (rotated difference + gadget decomposition) for the 32-bit
// accumulator, shared by the blind-rotation kernels.  cmux_head<AccT> (pbs_core.cuh) is the portable statement of the
// same computation and what the CPU emulators run.
#pragma once
#include "pbs_core.cuh"

namespace fsc {

// ---- head for the 32-bit accumulator, written for the ALU / FMA pipe split --------------------------------
// Same result as cmux_head<uint32_t> (pbs_core.cuh).  Over the 32 elements of a lane the rotated index crosses a
// multiple of 1024 at most once, so the swap / sign configuration of X^a takes two values per lane, A before the
// crossing and B after it; per element c in {0, 1} blends them with integer multiply-adds (FMA pipe) instead of
// predicated selects, and the int -> double conversion goes through the mantissa trick instead of the quarter-rate
// conversion unit.  About 13 ALU + 6 FMA-pipe + 2 FP64 instructions per coefficient pair, no predicates.
__device__ __forceinline__ int imad(int a, int b, int c) {
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void stream_head_u32(int lane, const pair_t<uint32_t>* poly, int a, int base_log, cplx (&z)[32]) {
    const int sh = 32 - base_log;
    const int half = 1 << (sh - 1);
    const int base = (lane - a) & 4095;
    const int q0 = base >> 10, q1 = (q0 + 1) & 3;
    // quadrant q: (x, y) <- 0: (+x, +y)  1: (+y, -x)  2: (-x, -y)  3: (-y, +x)
    const int swA = q0 & 1;
    const int sxA = 1 - (q0 & 2), syA = 1 - ((q0 ^ (q0 << 1)) & 2);
    const int sxB = 1 - (q1 & 2), syB = 1 - ((q1 ^ (q1 << 1)) & 2);
    const int dsx = sxB - sxA, dsy = syB - syA;
    unsigned b8 = (unsigned)(base & 1023) << 3;
    const char* pb = reinterpret_cast<const char*>(poly);
#pragma unroll
    for (int j2 = 0; j2 < 32; ++j2) {
        // ordering point at the start of every group of 4: the group's address arithmetic starts from a value the
        // compiler cannot see through, so it is not hoisted (and spilled) ahead of the previous groups
        if ((j2 & 3) == 0) asm volatile("" : "+r"(b8));
        const unsigned u = b8 + 256u * j2;                       // byte offset of the rotated pair, bit 13 = crossed
        const int c = (int)(u >> 13);
        const uint2 P = *reinterpret_cast<const uint2*>(pb + (u & 8191u));
        const uint2 O = *reinterpret_cast<const uint2*>(pb + lane * 8 + 256 * j2);
        const int sw = swA ^ c;
        const int sx = imad(c, dsx, sxA), sy = imad(c, dsy, syA);
        const int e = (int)(P.y - P.x);
        const int px = imad(sw, e, (int)P.x);
        const int py = (int)(P.x + P.y) - px;
        const int dx = imad(px, sx, half - (int)O.x);
        const int dy = imad(py, sy, half - (int)O.y);
#ifndef FSC_HEAD_MANTISSA      // int -> double through the conversion unit (A/B in one call, profiles/r02s_*: narrow levels -1.5 %, 149-296 blocks -5 %); FSC_HEAD_MANTISSA: the mantissa trick
        z[j2].x = (double)(dx >> sh);
        z[j2].y = (double)(dy >> sh);
#else
        z[j2].x = __hiloint2double(0x43300000, (dx >> sh) ^ (int)0x80000000) - 4503601774854144.0;
        z[j2].y = __hiloint2double(0x43300000, (dy >> sh) ^ (int)0x80000000) - 4503601774854144.0;
#endif
        // compiler-only ordering point every 4 elements: the four results must exist here, so the integer halves of
        // later elements cannot all be computed (and spilled) before the first conversion
        if ((j2 & 3) == 3) {
            asm volatile("" : "+d"(z[j2 - 3].x), "+d"(z[j2 - 3].y), "+d"(z[j2 - 2].x), "+d"(z[j2 - 2].y),
                              "+d"(z[j2 - 1].x), "+d"(z[j2 - 1].y), "+d"(z[j2].x), "+d"(z[j2].y) :: "memory");
        }
    }
}
template <typename AccT>
__device__ __forceinline__ void stream_head(int lane, const pair_t<AccT>* poly, int a, int base_log, cplx (&z)[32]) {
    if constexpr (sizeof(AccT) == 4) stream_head_u32(lane, poly, a, base_log, z);
    else cmux_head<AccT>(lane, poly, a, base_log, z);
}

}  // namespace fsc
