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


def test_default_keygen_draws_os_entropy():
    """the default key must not be re-derivable - two default clients differ, and each works."""
    a, b = ClientKey("toy"), ClientKey("toy")
    assert not np.array_equal(a.secret_keys()[0], b.secret_keys()[0])
    assert not np.array_equal(a.secret_keys()[1], b.secret_keys()[1])
    vals = np.arange(16, dtype=np.uint8)
    assert (a.decrypt_block_values(a.encrypt_block_values(vals)) == vals).all()
    assert 300 < int(a.secret_keys()[1].sum()) < 1748            # a binary key of 2048 bits, not a constant


def test_encryption_randomness_is_per_instance_even_with_the_same_key_seed():
    """encryption masks never come from (seed, counter): two instances with the same test seed (or one
    key file loaded twice) must not emit the same mask, or b1 - b2 would leak delta * (m1 - m2)."""
    a, b = ClientKey("toy", seed=9), ClientKey("toy", seed=9)
    vals = np.arange(8, dtype=np.uint8)
    ca, cb = a.encrypt_block_values(vals), b.encrypt_block_values(vals)
    assert not np.array_equal(ca[:, :2048], cb[:, :2048])
    assert (b.decrypt_block_values(ca) == vals).all()            # same secret key
    # the explicit test hook pins it (reproducible ciphertexts for multi-rank tests)
    c, d = ClientKey("toy", seed=9, encryption_seed=3), ClientKey("toy", seed=9, encryption_seed=3)
    assert np.array_equal(c.encrypt_block_values(vals), d.encrypt_block_values(vals))


def test_key_file_header_is_range_checked(ck, tmp_path):
    """base_log * level and the tuniform bound feed shifts - a crafted header is refused."""
    import struct
    from fhe_sign_b200.capi import FscError
    path = tmp_path / "client.fsc"
    ck.save(path)
    raw = bytearray(path.read_bytes())

    def refused(offset, value):
        bad = bytearray(raw)
        bad[offset:offset + 4] = struct.pack("<I", value)
        body = bytes(bad[:-8])
        h = 14695981039346656037
        for byte in body:
            h = ((h ^ byte) * 1099511628211) & (2**64 - 1)
        (tmp_path / "crafted.fsc").write_bytes(body + struct.pack("<Q", h))      # checksum fixed up: only the range check can refuse it
        with pytest.raises(FscError, match="header|parameter"):
            ClientKey.load(tmp_path / "crafted.fsc")

    refused(16 + 4 * 3, 64)          # pbs_base_log 64: the 64 - base_log * level shift would be undefined
    refused(16 + 4 * 4, 1 << 30)     # pbs_level huge: base_log * level overflows u32
    refused(56 + 4, 200)             # lwe_tuniform_bound (only read for the tuniform kind, still range-checked there)


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


# ---- on-disk formats (csrc/keyfile.cpp) ----

def test_client_key_file_roundtrip(ck, tmp_path):
    """the client key file re-derives the same secret and server keys and never replays encryption randomness."""
    from fhe_sign_b200.client import ClientKey
    vals = np.arange(16, dtype=np.uint8)
    ct_before = ck.encrypt_block_values(vals)
    path = tmp_path / "client.fsc"
    ck.save(path)
    assert path.stat().st_size < 64 * 1024                       # parameters + master key + secret bits, not 123 MB
    ck2 = ClientKey.load(path)
    assert ck2.params.lwe_dim == ck.params.lwe_dim and ck2.params.pbs_base_log == ck.params.pbs_base_log
    for a, b in zip(ck.secret_keys(), ck2.secret_keys()):
        assert np.array_equal(a, b)
    for a, b in zip(ck.server_keys(), ck2.server_keys()):
        assert np.array_equal(a, b)
    assert (ck2.decrypt_block_values(ct_before) == vals).all()   # old ciphertexts decrypt under the loaded key
    ct_after = ck2.encrypt_block_values(vals)
    assert (ck.decrypt_block_values(ct_after) == vals).all()
    assert not np.array_equal(ct_after, ct_before)               # fresh OS entropy per instance, nothing persisted
    ck3 = ClientKey.load(path)                                   # the same file loaded twice: still no shared masks
    assert not np.array_equal(ck3.encrypt_block_values(vals)[:, :2048], ct_after[:, :2048])


def test_server_key_and_block_files_roundtrip_and_reject_corruption(ck, tmp_path):
    from fhe_sign_b200.capi import FscError
    from fhe_sign_b200.client import load_blocks, load_server_keys, save_blocks
    path = tmp_path / "server.fsc"
    ck.save_server_keys(path)
    p, bsk, ksk = load_server_keys(path)
    assert p.lwe_dim == ck.params.lwe_dim and p.ks_level == ck.params.ks_level
    assert np.array_equal(bsk, ck.server_keys()[0]) and np.array_equal(ksk, ck.server_keys()[1])
    ct = ck.encrypt_block_values(np.arange(12, dtype=np.uint8))
    bpath = tmp_path / "ct.fsc"
    save_blocks(bpath, ck.params, ct)
    p2, ct2 = load_blocks(bpath)
    assert p2.poly_size == 2048 and np.array_equal(ct, ct2)
    save_blocks(tmp_path / "empty.fsc", ck.params, np.empty((0, 2049), np.uint64))
    assert load_blocks(tmp_path / "empty.fsc")[1].shape == (0, 2049)
    # a flipped payload byte, a truncated file, the wrong kind and a foreign file are all refused
    raw = bytearray(bpath.read_bytes())
    raw[200] ^= 1
    (tmp_path / "bad.fsc").write_bytes(bytes(raw))
    with pytest.raises(FscError, match="checksum"):
        load_blocks(tmp_path / "bad.fsc")
    (tmp_path / "short.fsc").write_bytes(bpath.read_bytes()[:-100])
    with pytest.raises(FscError):
        load_blocks(tmp_path / "short.fsc")
    with pytest.raises(FscError, match="kind"):
        load_server_keys(bpath)
    (tmp_path / "foreign.fsc").write_bytes(b"not a key file" * 20)
    with pytest.raises(FscError, match="FSCFILE1"):
        load_blocks(tmp_path / "foreign.fsc")
    with pytest.raises(FscError):
        load_blocks(tmp_path / "missing.fsc")
