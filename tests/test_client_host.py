"""Product client side (fsc_client_*, host CPU): round trips and key-material sanity against exact arithmetic
from the oracle (no GPU needed)."""
import numpy as np
import pytest

from fhe_sign_b200.client import ClientKey


@pytest.fixture(scope="module")
def ck():
    return ClientKey("toy", seed=5)


def test_encrypt_decrypt_roundtrip_and_noise(ck):
    vals = np.arange(16, dtype=np.uint8).repeat(4)
    ct = ck.encrypt_block_values(vals)
    got, noise = ck.decrypt_block_values(ct, with_noise=True)
    assert (got == vals).all()
    assert np.abs(noise).max() < 2**22                    # GLWE noise std 2^-48.3 of q = 2^64
    ct2 = ck.encrypt_block_values(vals)
    assert not np.array_equal(ct, ct2)                    # fresh randomness per call
    with pytest.raises(Exception):
        ck.encrypt_block_values([16])


def test_keygen_is_seeded(ck):
    a, b = ClientKey("toy", seed=5), ClientKey("toy", seed=6)
    assert np.array_equal(a.server_keys()[1], ck.server_keys()[1])
    assert not np.array_equal(b.server_keys()[1], ck.server_keys()[1])


def test_keyswitching_key_rows_encrypt_the_big_key_bits(ck):
    lwe, glwe = ck.secret_keys()
    _, ksk = ck.server_keys()
    n = ck.params.lwe_dim
    ksk = ksk.reshape(2048, 5, n + 1)
    for i in (0, 1, 777, 2047):
        for l in range(5):
            row = ksk[i, l]
            with np.errstate(over="ignore"):
                phase = (row[n] - (row[:n] * lwe).sum(dtype=np.uint64)).astype(np.uint64)
            want = np.uint64(int(glwe[i]) << (64 - 3 * (l + 1)))
            err = np.int64(phase - want)
            assert abs(int(err)) < 2**50, (i, l)            # LWE noise std 2^-18 of q


def test_bootstrapping_key_rows_encrypt_the_small_key_bits(ck, orc):
    lwe, glwe = ck.secret_keys()
    bsk, _ = ck.server_keys()
    n = ck.params.lwe_dim
    bsk = bsk.reshape(n, 2, 1, 2, 2048)
    S = glwe.astype(np.int64)
    fac = 1 << (64 - 23)
    for i in (0, 3, n - 1):
        for p in range(2):
            a, body = bsk[i, p, 0, 0], bsk[i, p, 0, 1]
            phase = (body - orc.negacyclic_mul_exact(a, S)).astype(np.uint64)
            want = np.zeros(2048, dtype=np.uint64)
            if lwe[i]:
                if p == 0:
                    want = (np.uint64(0) - glwe * np.uint64(fac)).astype(np.uint64)       # -s_i * S(X) * q/beta
                else:
                    want[0] = fac
            err = (phase - want).astype(np.int64)
            assert np.abs(err).max() < 2**22, (i, p)
