"""Python view of the radix-integer C ABI (fsc_radix_* in include/fhe_sign_cuda.h).

`RadixApi` is bound to a loaded library + context handle, so the same class drives the CUDA engine
(fhe_sign_b200.Context.radix) and the CPU mock of the circuit tests (tests/host/libfsc_mock.so)."""
import ctypes as C

import numpy as np

OPS = dict(add=0, sub=1, mul=2, min=3, max=4, shr=5, shl=6, and_=7, or_=8, xor=9, lt=10, eq=11, div=12, rem=13)
WORDS = 2049

RADIX_EXPORTS = ["fsc_radix_from_lwe", "fsc_radix_to_lwe", "fsc_radix_trivial", "fsc_radix_clone", "fsc_radix_free",
                 "fsc_radix_len", "fsc_radix_binary", "fsc_radix_scalar", "fsc_radix_mul_wide", "fsc_radix_mul_add_wide", "fsc_radix_scalar_mul_add_wide", "fsc_radix_cast",
                 "fsc_radix_slice", "fsc_radix_concat", "fsc_radix_sum", "fsc_radix_select", "fsc_radix_stats",
                 "fsc_radix_stats2", "fsc_set_level_exchange", "fsc_peer_pool_export", "fsc_peer_pool_connect",
                 "fsc_peer_pool_disconnect"]


def declare(L):
    vp, sz, u32 = C.c_void_p, C.c_size_t, C.c_uint32
    pp = C.POINTER(vp)
    sig = {
        "fsc_radix_from_lwe": [vp, vp, sz, pp], "fsc_radix_to_lwe": [vp, vp, vp],
        "fsc_radix_trivial": [vp, vp, sz, sz, pp], "fsc_radix_clone": [vp, vp, pp], "fsc_radix_free": [vp, vp],
        "fsc_radix_len": [vp, C.POINTER(sz)], "fsc_radix_binary": [vp, u32, vp, vp, pp],
        "fsc_radix_scalar": [vp, u32, vp, vp, sz, pp], "fsc_radix_mul_wide": [vp, vp, vp, sz, pp],
        "fsc_radix_mul_add_wide": [vp, vp, vp, vp, sz, pp],
        "fsc_radix_scalar_mul_add_wide": [vp, vp, vp, sz, vp, sz, pp],
        "fsc_radix_cast": [vp, vp, sz, pp], "fsc_radix_slice": [vp, vp, sz, sz, pp],
        "fsc_radix_concat": [vp, vp, sz, pp], "fsc_radix_sum": [vp, vp, sz, sz, pp],
        "fsc_radix_select": [vp, vp, vp, vp, pp], "fsc_radix_stats": [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
        "fsc_radix_stats2": [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
        "fsc_set_level_exchange": [vp, C.c_int32, C.c_int32, sz, vp, sz, vp, vp],
        "fsc_peer_pool_export": [vp, sz, vp], "fsc_peer_pool_connect": [vp, C.c_int32, C.c_int32, sz, vp],
        "fsc_peer_pool_disconnect": [vp],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int32


def int_to_le(v, n_bytes=None):
    v = int(v)
    n = max(1, (v.bit_length() + 7) // 8) if n_bytes is None else n_bytes
    return (C.c_uint8 * n).from_buffer_copy(v.to_bytes(n, "little")), n


class RadixValue:
    """Handle of one device-resident radix integer (e.g. an FheUint32 = 16 blocks)."""

    def __init__(self, api, h):
        self.api, self.h = api, h

    def __del__(self):
        try:
            if self.h and self.api.alive():
                self.api.L.fsc_radix_free(self.api.ctx, self.h)
            self.h = None
        except Exception:
            pass

    def __len__(self):
        n = C.c_size_t()
        self.api.L.fsc_radix_len(self.h, C.byref(n))
        return n.value

    # operator sugar mirroring the reference's use of tfhe's overloaded operators
    def __add__(self, o): return self.api.binary("add", self, o) if isinstance(o, RadixValue) else self.api.scalar("add", self, o)
    def __sub__(self, o): return self.api.binary("sub", self, o)
    def __mul__(self, o): return self.api.binary("mul", self, o) if isinstance(o, RadixValue) else self.api.scalar("mul", self, o)
    def __rshift__(self, o): return self.api.binary("shr", self, o) if isinstance(o, RadixValue) else self.api.scalar("shr", self, o)
    def __lshift__(self, o): return self.api.binary("shl", self, o) if isinstance(o, RadixValue) else self.api.scalar("shl", self, o)
    def __and__(self, o): return self.api.binary("and_", self, o) if isinstance(o, RadixValue) else self.api.scalar("and_", self, o)
    def __floordiv__(self, o): return self.api.scalar("div", self, o)
    def __mod__(self, o): return self.api.scalar("rem", self, o)


class RadixApi:
    def __init__(self, L, ctx, check, alive=lambda: True):
        self.L, self.ctx, self._check, self.alive = L, ctx, check, alive
        declare(L)

    def _new(self, fn, *args):
        h = C.c_void_p()
        self._check(fn(self.ctx, *args, C.byref(h)))
        return RadixValue(self, h)

    def from_lwe(self, blocks):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint64).reshape(-1, WORDS)
        return self._new(self.L.fsc_radix_from_lwe, blocks.ctypes.data_as(C.c_void_p), blocks.shape[0])

    def to_lwe(self, r):
        out = np.empty((len(r), WORDS), dtype=np.uint64)
        self._check(self.L.fsc_radix_to_lwe(self.ctx, r.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def trivial(self, value, n_blocks):
        buf, n = int_to_le(value)
        return self._new(self.L.fsc_radix_trivial, buf, n, n_blocks)

    def clone(self, a):
        return self._new(self.L.fsc_radix_clone, a.h)

    def binary(self, op, a, b):
        return self._new(self.L.fsc_radix_binary, OPS[op], a.h, b.h)

    def scalar(self, op, a, scalar):
        buf, n = int_to_le(scalar)
        return self._new(self.L.fsc_radix_scalar, OPS[op], a.h, buf, n)

    def mul_wide(self, a, b, out_blocks):
        return self._new(self.L.fsc_radix_mul_wide, a.h, b.h, out_blocks)

    def mul_add_wide(self, a, b, addend, out_blocks):
        """a * b + addend with one carry propagation (the addend joins the product's column sum)."""
        return self._new(self.L.fsc_radix_mul_add_wide, a.h, b.h, addend.h, out_blocks)

    def scalar_mul_add_wide(self, a, scalar, addend, out_blocks):
        """a * scalar + addend (scalar in plaintext), one carry propagation."""
        buf, n = int_to_le(scalar)
        return self._new(self.L.fsc_radix_scalar_mul_add_wide, a.h, buf, n, addend.h if addend is not None else None, out_blocks)

    def cast(self, a, n_blocks):
        return self._new(self.L.fsc_radix_cast, a.h, n_blocks)

    def slice(self, a, first, n_blocks):
        return self._new(self.L.fsc_radix_slice, a.h, first, n_blocks)

    def concat(self, parts):
        arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
        return self._new(self.L.fsc_radix_concat, arr, len(parts))

    def sum(self, operands, n_blocks):
        arr = (C.c_void_p * len(operands))(*[p.h for p in operands])
        return self._new(self.L.fsc_radix_sum, arr, len(operands), n_blocks)

    def select(self, cond, a, b):
        return self._new(self.L.fsc_radix_select, cond.h, a.h, b.h)

    def min(self, a, b): return self.binary("min", a, b)
    def max(self, a, b): return self.binary("max", a, b)
    def lt(self, a, b): return self.binary("lt", a, b)
    def eq(self, a, b): return self.binary("eq", a, b)

    def sharded_levels(self):
        n = C.c_uint64()
        self.L.fsc_radix_stats2(self.ctx, None, None, C.byref(n))
        return n.value

    def stats(self):
        p, l = C.c_uint64(), C.c_uint64()
        self.L.fsc_radix_stats(self.ctx, C.byref(p), C.byref(l))
        return p.value, l.value
