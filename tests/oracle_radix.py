"""Test helper / CPU baseline: the radix C ABI running on the CPU oracle (tests/host/oracle_backend.cpp).

Every block is a real LWE ciphertext under the oracle's seeded keys; a level is keyswitch + PBS on the host cores
(OpenMP).  Same RadixApi as the GPU context, so BigUintFHE / sign_fhe_with_k0 and every operator run unchanged on it.
Test infrastructure: only tests/ and bench.py's CPU legs import this."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


class OracleRadix:
    def __init__(self, K, threads=0):
        """K: oracle.orc.Keys; threads: OpenMP threads per level (0 = omp default)."""
        subprocess.check_call(["make", "-C", os.path.join(HERE, "host"), "-s", "libfsc_orc.so"])
        from oracle import orc
        orc.lib()                                                    # liborc.so first, so that the backend binds to the same copy
        self.L = C.CDLL(os.path.join(HERE, "host", "libfsc_orc.so"))
        from fhe_sign_b200.radix import RadixApi
        self.L.fscorc_ctx_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        self.L.fscorc_last_error.argtypes = [C.c_void_p]
        self.L.fscorc_last_error.restype = C.c_char_p
        self.L.fscorc_ctx_destroy.argtypes = [C.c_void_p]
        self.K = K
        h = C.c_void_p()
        rc = self.L.fscorc_ctx_create(K._h, threads, C.byref(h))
        assert rc == 0
        self.ctx = h
        self.api = RadixApi(self.L, h, self._check, alive=lambda: bool(self.ctx))      # values outliving close() must not touch the freed context

    @property
    def radix(self):
        return self.api

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("radix error %d: %s" % (rc, self.L.fscorc_last_error(self.ctx).decode()))

    def sync(self):
        pass

    def close(self):
        if self.ctx:
            self.L.fscorc_ctx_destroy(self.ctx)
            self.ctx = None
