"""Client side (host CPU) of the path over the fsc_client_* C ABI: the ClientKey look-alike.

Mirrors what the reference gets from tfhe: `generate_keys(config)` (src/biguint.rs:277) -> (ClientKey, server key
material), `FheUint32::try_encrypt(value, &client_key)` (src/biguint.rs:26) and `.decrypt(&client_key)`
(src/biguint.rs:70)."""
import ctypes as C

import numpy as np

from .capi import FscError, Params, load_library

WORDS = 2049

NOISE = {
    "2_2_gaussian": dict(noise_kind=0, lwe_noise_std=3.5539902359442825e-06, glwe_noise_std=2.845267479601915e-15),
    "2_2_tuniform": dict(noise_kind=1, lwe_tuniform_bound=46, glwe_tuniform_bound=17),
    "toy": dict(noise_kind=0, lwe_noise_std=3.5539902359442825e-06, glwe_noise_std=2.845267479601915e-15),
}


class NoiseParams(C.Structure):
    _fields_ = [("noise_kind", C.c_uint32), ("lwe_tuniform_bound", C.c_uint32), ("glwe_tuniform_bound", C.c_uint32),
                ("reserved", C.c_uint32), ("lwe_noise_std", C.c_double), ("glwe_noise_std", C.c_double)]


CLIENT_EXPORTS = ["fsc_client_keygen", "fsc_client_keygen_seeded", "fsc_client_set_encryption_seed", "fsc_client_free", "fsc_client_last_error", "fsc_client_server_keys",
                  "fsc_client_secret_keys", "fsc_client_encrypt_blocks", "fsc_client_decrypt_blocks",
                  "fsc_client_save", "fsc_client_load", "fsc_server_keys_save", "fsc_server_keys_load",
                  "fsc_blocks_save", "fsc_blocks_load", "fsc_buffer_free"]


def _declare(L):
    vp, sz = C.c_void_p, C.c_size_t
    L.fsc_client_keygen.argtypes = [C.POINTER(Params), C.POINTER(NoiseParams), C.POINTER(vp)]
    L.fsc_client_keygen_seeded.argtypes = [C.POINTER(Params), C.POINTER(NoiseParams), C.c_uint64, C.POINTER(vp)]
    L.fsc_client_set_encryption_seed.argtypes = [vp, C.c_uint64]
    L.fsc_client_free.argtypes = [vp]
    L.fsc_client_last_error.argtypes = [vp]; L.fsc_client_last_error.restype = C.c_char_p
    L.fsc_client_server_keys.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz)]
    L.fsc_client_secret_keys.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.fsc_client_encrypt_blocks.argtypes = [vp, vp, sz, vp]
    L.fsc_client_decrypt_blocks.argtypes = [vp, vp, sz, vp, vp]
    L.fsc_client_save.argtypes = [vp, C.c_char_p]
    L.fsc_client_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.fsc_server_keys_save.argtypes = [vp, C.c_char_p]
    L.fsc_server_keys_load.argtypes = [C.c_char_p, C.POINTER(Params), C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz)]
    L.fsc_blocks_save.argtypes = [C.c_char_p, C.POINTER(Params), vp, sz]
    L.fsc_blocks_load.argtypes = [C.c_char_p, C.POINTER(Params), C.POINTER(vp), C.POINTER(sz)]
    L.fsc_buffer_free.argtypes = [vp]


class ClientKey:
    """Secret keys + the server key material derived from them.

    seed=None (default): 256-bit master key from the OS CSPRNG, like the reference's tfhe::generate_keys.
    seed=<int>: TEST-ONLY deterministic keys (at most 64 bits of entropy).  Encryption masks and noise always come from a
    fresh OS-entropy stream per instance unless `encryption_seed` pins them too (reproducible ciphertexts in tests)."""

    def __init__(self, preset="2_2_gaussian", seed=None, encryption_seed=None, _handle=None, _params=None):
        self.L = load_library()
        _declare(self.L)
        if _handle is not None:      # ClientKey.load
            self.h, self.params, self.noise = _handle, _params, None
            return
        self.params = Params.preset(preset)
        self.noise = NoiseParams(**NOISE[preset])
        h = C.c_void_p()
        if seed is None:
            rc = self.L.fsc_client_keygen(C.byref(self.params), C.byref(self.noise), C.byref(h))
        else:
            rc = self.L.fsc_client_keygen_seeded(C.byref(self.params), C.byref(self.noise), int(seed), C.byref(h))
        if rc != 0:
            raise FscError(rc, (self.L.fsc_client_last_error(None) or b"").decode())
        self.h = h
        if encryption_seed is not None:
            self._check(self.L.fsc_client_set_encryption_seed(self.h, int(encryption_seed)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.fsc_client_free(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise FscError(rc, (self.L.fsc_client_last_error(self.h) or b"").decode())

    # ---- on-disk formats (csrc/keyfile.cpp) ---------------------------------------------------------------
    def save(self, path):
        """client key file: parameters, noise, 256-bit master key, secret bits (a few KB; server keys are re-derived on
        load; no encryption state is stored).  Holds the secret key: protect it accordingly."""
        self._check(self.L.fsc_client_save(self.h, str(path).encode()))

    @classmethod
    def load(cls, path):
        L = load_library()
        _declare(L)
        h = C.c_void_p()
        rc = L.fsc_client_load(str(path).encode(), C.byref(h))
        if rc != 0:
            raise FscError(rc, (L.fsc_client_last_error(None) or b"").decode())
        return cls(_handle=h, _params=_file_params(path))

    def save_server_keys(self, path):
        """expanded bootstrapping + keyswitching keys, no secrets: what the GPU host needs for fsc_keys_upload."""
        self._check(self.L.fsc_server_keys_save(self.h, str(path).encode()))

    def server_keys(self):
        """(bsk_std, ksk) as numpy views, ready for Context.upload_keys."""
        b, k, nb, nk = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._check(self.L.fsc_client_server_keys(self.h, C.byref(b), C.byref(nb), C.byref(k), C.byref(nk)))
        bsk = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint64)), shape=(nb.value,))
        ksk = np.ctypeslib.as_array(C.cast(k, C.POINTER(C.c_uint64)), shape=(nk.value,))
        return bsk, ksk

    def secret_keys(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self.L.fsc_client_secret_keys(self.h, C.byref(a), C.byref(b)))
        lwe = np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_uint64)), shape=(self.params.lwe_dim,))
        glwe = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint64)), shape=(self.params.glwe_dim * self.params.poly_size,))
        return lwe, glwe

    def encrypt_block_values(self, values):
        values = np.ascontiguousarray(values, dtype=np.uint8)
        out = np.empty((values.size, WORDS), dtype=np.uint64)
        self._check(self.L.fsc_client_encrypt_blocks(self.h, values.ctypes.data_as(C.c_void_p), values.size, out.ctypes.data_as(C.c_void_p)))
        return out

    def decrypt_block_values(self, blocks, with_noise=False):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint64).reshape(-1, WORDS)
        vals = np.empty(blocks.shape[0], dtype=np.uint8)
        noise = np.empty(blocks.shape[0], dtype=np.int64) if with_noise else None
        self._check(self.L.fsc_client_decrypt_blocks(self.h, blocks.ctypes.data_as(C.c_void_p), blocks.shape[0], vals.ctypes.data_as(C.c_void_p),
                                                     None if noise is None else noise.ctypes.data_as(C.c_void_p)))
        return (vals, noise) if with_noise else vals

    # ---- the surface BigUintFHE uses (FheUint32::try_encrypt / decrypt) ---------------------------------
    def encrypt_blocks(self, value, n_blocks, api):
        digits = [(int(value) >> (2 * i)) & 3 for i in range(n_blocks)]
        return api.from_lwe(self.encrypt_block_values(digits))

    def encrypt_u32(self, value, api):
        return self.encrypt_blocks(int(value) & 0xFFFFFFFF, 16, api)

    def decrypt(self, r, api):
        d = self.decrypt_block_values(api.to_lwe(r))
        if (d >= 4).any():
            raise ValueError("decrypted block carries are not empty: %s" % d)
        return sum(int(v) << (2 * i) for i, v in enumerate(d))


def _file_params(path):
    """fsc_params stored in an FSCFILE1 container (offset 16, ten u32)."""
    raw = np.fromfile(str(path), dtype=np.uint32, count=14)[4:14]
    return Params(*[int(v) for v in raw])


def load_server_keys(path):
    """-> (Params, bsk_std, ksk) from a file written by ClientKey.save_server_keys (copies into numpy arrays)."""
    L = load_library()
    _declare(L)
    p, b, k, nb, nk = Params(), C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
    rc = L.fsc_server_keys_load(str(path).encode(), C.byref(p), C.byref(b), C.byref(nb), C.byref(k), C.byref(nk))
    if rc != 0:
        raise FscError(rc, (L.fsc_client_last_error(None) or b"").decode())
    try:
        bsk = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint64)), shape=(nb.value,)).copy()
        ksk = np.ctypeslib.as_array(C.cast(k, C.POINTER(C.c_uint64)), shape=(nk.value,)).copy()
    finally:
        L.fsc_buffer_free(b)
    return p, bsk, ksk


def save_blocks(path, params, blocks):
    L = load_library()
    _declare(L)
    blocks = np.ascontiguousarray(blocks, dtype=np.uint64).reshape(-1, params.glwe_dim * params.poly_size + 1)
    rc = L.fsc_blocks_save(str(path).encode(), C.byref(params), blocks.ctypes.data_as(C.c_void_p), blocks.shape[0])
    if rc != 0:
        raise FscError(rc, (L.fsc_client_last_error(None) or b"").decode())


def load_blocks(path):
    L = load_library()
    _declare(L)
    p, b, n = Params(), C.c_void_p(), C.c_size_t()
    rc = L.fsc_blocks_load(str(path).encode(), C.byref(p), C.byref(b), C.byref(n))
    if rc != 0:
        raise FscError(rc, (L.fsc_client_last_error(None) or b"").decode())
    try:
        words = p.glwe_dim * p.poly_size + 1
        out = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint64)), shape=(n.value, words)).copy() if n.value else np.empty((0, words), np.uint64)
    finally:
        L.fsc_buffer_free(b)
    return p, out


def generate_keys(preset="2_2_gaussian", seed=None, encryption_seed=None):
    """tfhe::generate_keys look-alike: returns (client_key, (bsk_std, ksk)).  seed=None: OS entropy (the default, as in
    the reference); an integer seed gives reproducible keys for tests and benches only."""
    ck = ClientKey(preset, seed, encryption_seed)
    return ck, ck.server_keys()
