// pbs_stream_tables.cuh — constant tables of the stream formulation (pbs_core2.cuh) as one global-memory image that
// every CTA of pbs_stream_kernel / pbs_split_kernel copies into shared memory.
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <vector>
#include "pbs_core2.cuh"
#include "fsc_internal.h"

namespace fsc {

// ---- constant tables (global memory image, copied into shared memory by every CTA) --------------------
// [0, 16)        pass 0 (forward 1, uniform)          [16, 32)      pass 2 (inverse A, uniform)
// [32, 544)      pass 1 (forward 2, [ci][lane])       [544, 1056)   pass 3 (inverse B, [ci][lane])
// [1056, 2080)   twist [pos][lane]
constexpr int kTabU0 = 0, kTabU2 = 16, kTabL1 = 32, kTabL3 = 544, kTabTwist = 1056, kTabCplx = 2080;

template <typename AccT>
const cplx* stream_tables() {      // device pointer, built once per device and accumulator type
    static cplx* per_dev[64] = {};
    static std::mutex mu;      // contexts of different host threads may reach this together
    int dev = 0;
    FSC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    cplx*& d = per_dev[dev & 63];
    if (!d) {
        std::vector<cplx> h(kTabCplx);
        for (int ci = 0; ci < 16; ++ci) {
            h[kTabU0 + ci] = pass_const(ci, pass_g(0, 0));
            h[kTabU2 + ci] = pass_const(ci, pass_g(2, 0));
            for (int l = 0; l < 32; ++l) {
                h[kTabL1 + ci * 32 + l] = pass_const(ci, pass_g(1, l));
                h[kTabL3 + ci * 32 + l] = pass_const(ci, pass_g(3, l));
            }
        }
        for (int pos = 0; pos < 32; ++pos)
            for (int l = 0; l < 32; ++l) h[kTabTwist + pos * 32 + l] = twist_const<AccT>(pos, l);
        FSC_CUDA_CHECK(cudaMalloc(&d, kTabCplx * sizeof(cplx)));
        FSC_CUDA_CHECK(cudaMemcpy(d, h.data(), kTabCplx * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    return d;
}

__device__ __forceinline__ StridedConsts pass_table(const cplx* tabs, int q, int lane) {
    // q = 0, 2: uniform tables at 0 and 16;  q = 1, 3: per-lane tables at 32 and 544
    StridedConsts c;
    c.base = tabs + ((q & 1) ? (q == 1 ? kTabL1 : kTabL3) + lane : (q == 0 ? kTabU0 : kTabU2));
    c.stride = (q & 1) ? 32 : 1;
    return c;
}

}  // namespace fsc
