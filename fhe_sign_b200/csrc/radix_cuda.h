// radix_cuda.h — the CUDA side of the radix layer (device block pool + level execution).
#pragma once
#include "radix.h"

namespace fsc {
struct Engine;
RadixBackend* make_cuda_backend(Engine* eng);
}
