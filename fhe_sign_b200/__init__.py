"""fhe_sign_b200 — B200-native (sm_100a) TFHE bootstrapping engine for fhe-sign's BigUintFHE path.

The compute lives in csrc/ (hand-written CUDA behind the C ABI of include/fhe_sign_cuda.h).  This
package is the thin Python host binding used by the tests and bench.py; it never falls back to a
CPU implementation: importing `fhe_sign_b200.capi` raises if libfhe_sign_cuda.so is missing, and
creating a context raises if no B200 is visible.
"""
from .capi import FscError, Params, Context, lib_path, load_library  # noqa: F401

__all__ = ["FscError", "Params", "Context", "lib_path", "load_library"]
