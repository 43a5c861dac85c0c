"""ctypes binding of libfhe_sign_cuda.so (include/fhe_sign_cuda.h) — no torch types at the boundary."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS = {0: "OK", 1: "BAD_ARG", 2: "PARAMS", 3: "OOM", 4: "CUDA", 5: "NO_KEYS", 6: "COMM", 7: "INTERNAL"}
LWE_BIG, LWE_SMALL = 0, 1


class FscError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fsc error %s (%d): %s" % (STATUS.get(code, "?"), code, msg))
        self.code = code


class Params(C.Structure):
    """Mirror of fsc_params (replaces tfhe ConfigBuilder::default(), src/biguint.rs:276)."""
    _fields_ = [("lwe_dim", C.c_uint32), ("glwe_dim", C.c_uint32), ("poly_size", C.c_uint32),
                ("pbs_base_log", C.c_uint32), ("pbs_level", C.c_uint32), ("ks_base_log", C.c_uint32),
                ("ks_level", C.c_uint32), ("message_modulus", C.c_uint32), ("carry_modulus", C.c_uint32),
                ("acc_bits", C.c_uint32)]

    PRESETS = {
        # PARAM_MESSAGE_2_CARRY_2_KS_PBS, Gaussian noise flavour (n = 834) and TUniform flavour (n = 887)
        "2_2_gaussian": dict(lwe_dim=834, pbs_base_log=23),
        "2_2_tuniform": dict(lwe_dim=887, pbs_base_log=22),
        "toy": dict(lwe_dim=48, pbs_base_log=23),
    }

    @classmethod
    def preset(cls, name, acc_bits=32):
        d = dict(glwe_dim=1, poly_size=2048, pbs_level=1, ks_base_log=3, ks_level=5,
                 message_modulus=4, carry_modulus=4, acc_bits=acc_bits)
        d.update(cls.PRESETS[name])
        return cls(**d)


def lib_path():
    return os.path.join(_HERE, "lib", "libfhe_sign_cuda.so")


def load_library():
    """Load the CUDA library; fails loudly when it has not been built (there is no CPU fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError("libfhe_sign_cuda.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                          "fhe_sign_b200 has no CPU fallback")
    L = C.CDLL(path)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32
    pp = C.POINTER(vp)
    sig = {
        "fsc_ctx_create": [C.POINTER(Params), i32, C.c_size_t, pp],
        "fsc_ctx_destroy": [vp],
        "fsc_get_params": [vp, C.POINTER(Params)],
        "fsc_sync": [vp],
        "fsc_keys_upload": [vp, vp, sz, vp, sz],
        "fsc_lwe_alloc": [vp, u32, sz, pp],
        "fsc_lwe_free": [vp, vp],
        "fsc_lwe_upload": [vp, vp, sz, vp, sz],
        "fsc_lwe_download": [vp, vp, sz, vp, sz],
        "fsc_lwe_info": [vp, C.POINTER(u32), C.POINTER(sz), C.POINTER(sz), pp],
        "fsc_luts_from_tables": [vp, vp, sz, pp],
        "fsc_luts_upload": [vp, vp, sz, pp],
        "fsc_luts_free": [vp, vp],
        "fsc_keyswitch_batch": [vp, vp, sz, vp, sz, sz],
        "fsc_pbs_batch": [vp, vp, sz, vp, vp, vp, sz, sz],
        "fsc_ks_pbs_batch": [vp, vp, sz, vp, vp, vp, sz, sz],
        "fsc_apply_lut_host": [vp, vp, vp, vp, vp, sz],
        "fsc_timer_start": [vp],
        "fsc_timer_stop": [vp, C.POINTER(C.c_float)],
        "fsc_launch_count": [vp, C.POINTER(C.c_uint64)],
        "fsc_debug_negacyclic_mul": [vp, vp, vp, vp, sz],
        "fsc_debug_fourier_key": [vp, vp, i32],
        "fsc_measure_fp64_peak": [vp, C.POINTER(C.c_double)],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = i32
    L.fsc_last_error.argtypes = [vp]
    L.fsc_last_error.restype = C.c_char_p
    L.fsc_pbs_kernel_name.argtypes = [vp]
    L.fsc_pbs_kernel_name.restype = C.c_char_p
    _LIB = L
    return L


EXPORTS = ["fsc_ctx_create", "fsc_ctx_destroy", "fsc_last_error", "fsc_get_params", "fsc_sync", "fsc_keys_upload",
           "fsc_lwe_alloc", "fsc_lwe_free", "fsc_lwe_upload", "fsc_lwe_download", "fsc_lwe_info",
           "fsc_luts_from_tables", "fsc_luts_upload", "fsc_luts_free", "fsc_keyswitch_batch", "fsc_pbs_batch",
           "fsc_ks_pbs_batch", "fsc_apply_lut_host", "fsc_timer_start", "fsc_timer_stop", "fsc_launch_count",
           "fsc_debug_negacyclic_mul", "fsc_debug_fourier_key", "fsc_measure_fp64_peak", "fsc_pbs_kernel_name"]
from .radix import RADIX_EXPORTS  # noqa: E402
EXPORTS = EXPORTS + RADIX_EXPORTS + ["fsc_client_keygen", "fsc_client_keygen_seeded", "fsc_client_set_encryption_seed", "fsc_client_free", "fsc_client_last_error", "fsc_client_server_keys",
                                     "fsc_client_secret_keys", "fsc_client_encrypt_blocks", "fsc_client_decrypt_blocks",
                                     "fsc_client_save", "fsc_client_load", "fsc_server_keys_save", "fsc_server_keys_load",
                                     "fsc_blocks_save", "fsc_blocks_load", "fsc_buffer_free"]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class LweArray:
    def __init__(self, ctx, kind, count):
        self.ctx, self.kind, self.count = ctx, kind, count
        h = C.c_void_p()
        ctx._check(ctx.L.fsc_lwe_alloc(ctx.h, kind, count, C.byref(h)))
        self.h = h
        self.words = (ctx.params.glwe_dim * ctx.params.poly_size + 1) if kind == LWE_BIG else ctx.params.lwe_dim + 1

    def upload(self, host, first=0):
        host = np.ascontiguousarray(host, dtype=np.uint64).reshape(-1, self.words)
        self.ctx._check(self.ctx.L.fsc_lwe_upload(self.ctx.h, self.h, first, _ptr(host), host.shape[0]))
        return self

    def download(self, first=0, count=None):
        count = self.count - first if count is None else count
        out = np.empty((count, self.words), dtype=np.uint64)
        self.ctx._check(self.ctx.L.fsc_lwe_download(self.ctx.h, self.h, first, _ptr(out), count))
        return out

    def device_ptr(self):
        p = C.c_void_p()
        self.ctx.L.fsc_lwe_info(self.h, None, None, None, C.byref(p))
        return p.value

    def free(self):
        if self.h:
            self.ctx.L.fsc_lwe_free(self.ctx.h, self.h)
            self.h = None


class Luts:
    def __init__(self, ctx, h, n):
        self.ctx, self.h, self.n = ctx, h, n

    def free(self):
        if self.h:
            self.ctx.L.fsc_luts_free(self.ctx.h, self.h)
            self.h = None


class Context:
    """One GPU server key (replaces the thread-local key installed by tfhe::set_server_key)."""

    def __init__(self, params, device=0, stream=0):
        self.L = load_library()
        self.params = params
        h = C.c_void_p()
        rc = self.L.fsc_ctx_create(C.byref(params), device, stream, C.byref(h))
        if rc != 0:
            raise FscError(rc, (self.L.fsc_last_error(None) or b"").decode())
        self.h = h

    def _check(self, rc):
        if rc != 0:
            raise FscError(rc, (self.L.fsc_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.fsc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.L.fsc_sync(self.h))

    @property
    def radix(self):
        """Radix-integer operator surface (FheUint8/32/64 look-alikes); available once keys are uploaded."""
        if getattr(self, "_radix", None) is None:
            from .radix import RadixApi
            self._radix = RadixApi(self.L, self.h, self._check, alive=lambda: bool(getattr(self, "h", None)))
        return self._radix

    def upload_keys(self, bsk_std, ksk):
        bsk_std = np.ascontiguousarray(bsk_std, dtype=np.uint64)
        ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
        self._check(self.L.fsc_keys_upload(self.h, _ptr(bsk_std), bsk_std.size, _ptr(ksk), ksk.size))

    def lwe(self, kind, count):
        return LweArray(self, kind, count)

    def luts_from_tables(self, tables):
        tables = np.ascontiguousarray(tables, dtype=np.uint64).reshape(-1, self.params.message_modulus * self.params.carry_modulus)
        h = C.c_void_p()
        self._check(self.L.fsc_luts_from_tables(self.h, _ptr(tables), tables.shape[0], C.byref(h)))
        return Luts(self, h, tables.shape[0])

    def luts_upload(self, polys):
        polys = np.ascontiguousarray(polys, dtype=np.uint64).reshape(-1, self.params.poly_size)
        h = C.c_void_p()
        self._check(self.L.fsc_luts_upload(self.h, _ptr(polys), polys.shape[0], C.byref(h)))
        return Luts(self, h, polys.shape[0])

    @staticmethod
    def _idx(lut_idx, count):
        if lut_idx is None:
            return None
        a = np.ascontiguousarray(lut_idx, dtype=np.uint32)
        assert a.size == count
        return a

    def keyswitch(self, in_big, out_small, count=None, in_first=0, out_first=0):
        count = in_big.count - in_first if count is None else count
        self._check(self.L.fsc_keyswitch_batch(self.h, in_big.h, in_first, out_small.h, out_first, count))

    def pbs(self, in_small, luts, lut_idx, out_big, count=None, in_first=0, out_first=0):
        count = in_small.count - in_first if count is None else count
        idx = self._idx(lut_idx, count)
        self._check(self.L.fsc_pbs_batch(self.h, in_small.h, in_first, luts.h, _ptr(idx), out_big.h, out_first, count))

    def ks_pbs(self, in_big, luts, lut_idx, out_big, count=None, in_first=0, out_first=0):
        count = in_big.count - in_first if count is None else count
        idx = self._idx(lut_idx, count)
        self._check(self.L.fsc_ks_pbs_batch(self.h, in_big.h, in_first, luts.h, _ptr(idx), out_big.h, out_first, count))

    def apply_lut_host(self, in_big_host, luts, lut_idx=None, out=None):
        """End-to-end call with host buffers: H2D, keyswitch + PBS, D2H."""
        words = self.params.glwe_dim * self.params.poly_size + 1
        assert in_big_host.dtype == np.uint64 and in_big_host.flags.c_contiguous
        count = in_big_host.size // words
        if out is None:
            out = np.empty((count, words), dtype=np.uint64)
        idx = self._idx(lut_idx, count)
        self._check(self.L.fsc_apply_lut_host(self.h, _ptr(in_big_host), luts.h, _ptr(idx), _ptr(out), count))
        return out

    def timer_start(self):
        self._check(self.L.fsc_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.L.fsc_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        self.L.fsc_launch_count(self.h, C.byref(n))
        return n.value

    def pbs_kernel_name(self):
        return self.L.fsc_pbs_kernel_name(self.h).decode()

    def measure_fp64_peak(self):
        t = C.c_double()
        self._check(self.L.fsc_measure_fp64_peak(self.h, C.byref(t)))
        return t.value

    def debug_fourier_key(self, stream_order=False):
        """the Fourier bootstrapping key held by the context: [n][32][4][32] complex (test hook)"""
        out = np.empty((self.params.lwe_dim, 32, 4, 32), dtype=np.complex128)
        self._check(self.L.fsc_debug_fourier_key(self.h, _ptr(out), 1 if stream_order else 0))
        return out

    def debug_negacyclic_mul(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, self.params.poly_size)
        b = np.ascontiguousarray(b, dtype=np.int64).reshape(-1, self.params.poly_size)
        c = np.empty_like(a)
        self._check(self.L.fsc_debug_negacyclic_mul(self.h, _ptr(a), _ptr(b), _ptr(c), a.shape[0]))
        return c
