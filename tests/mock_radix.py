"""Test helper: the radix C ABI running on the CPU mock backend (plaintext 'ciphertexts')."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DELTA = 1 << 59


class MockRadix:
    def __init__(self):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "host"), "-s"])
        self.L = C.CDLL(os.path.join(HERE, "host", "libfsc_mock.so"))
        from fhe_sign_b200.radix import RadixApi
        self.L.fscmock_ctx_create.argtypes = [C.POINTER(C.c_void_p)]
        self.L.fscmock_last_error.argtypes = [C.c_void_p]
        self.L.fscmock_last_error.restype = C.c_char_p
        self.L.fscmock_counters.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3
        self.L.fscmock_ctx_destroy.argtypes = [C.c_void_p]
        h = C.c_void_p()
        self.L.fscmock_ctx_create(C.byref(h))
        self.ctx = h
        self.api = RadixApi(self.L, h, self._check)

    @property
    def radix(self):
        return self.api

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("radix error %d: %s" % (rc, self.L.fscmock_last_error(self.ctx).decode()))

    # trivial LWE "encryption": zero mask, body = digit * delta
    def enc(self, value, n_blocks):
        ct = np.zeros((n_blocks, 2049), dtype=np.uint64)
        for i in range(n_blocks):
            ct[i, 2048] = ((int(value) >> (2 * i)) & 3) * DELTA
        return self.api.from_lwe(ct)

    def dec(self, r):
        ct = self.api.to_lwe(r)
        v = 0
        for i in range(ct.shape[0]):
            d = int((int(ct[i, 2048]) + DELTA // 2) // DELTA) % 32
            assert d < 4, "block %d is not clean: %d" % (i, d)
            v |= d << (2 * i)
        return v

    def counters(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.L.fscmock_counters(self.ctx, C.byref(a), C.byref(b), C.byref(c))
        return dict(violations=a.value, live_slots=b.value, max_batch=c.value)


class MockClientKey:
    """ClientKey stand-in for the mock backend: 'encrypts' u32 digits as trivial LWE blocks."""

    def __init__(self, mock):
        self.m = mock

    def encrypt_u32(self, value, api):
        return self.m.enc(value, 16)

    def decrypt(self, r, api):
        return self.m.dec(r)
