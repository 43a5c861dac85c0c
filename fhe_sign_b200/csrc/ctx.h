// ctx.h — the object behind the opaque fsc_ctx / fsc_radix handles of the C ABI.
#pragma once
#include <string>

#include "../../include/fhe_sign_cuda.h"
#include "radix.h"

namespace fsc {
struct Engine;
}

struct fsc_ctx {
    fsc::Engine* eng = nullptr;          // device state (null in the CPU mock used by the circuit tests)
    fsc::RadixBackend* rb = nullptr;     // radix backend (CUDA pool or mock)
    fsc::Evaluator* ev = nullptr;
    fsc_params params{};
    std::string err;
};

struct fsc_radix {
    fsc::Radix blocks;
};
