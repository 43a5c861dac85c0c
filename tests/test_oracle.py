"""CPU tests of the oracle itself (no GPU): primitives against exact arithmetic and the plaintext-level
known answers the reference pins."""
import numpy as np
import pytest


def test_decompose_reconstructs(orc, rng):
    for base_log, level in ((3, 5), (23, 1), (22, 1), (4, 3)):
        for x in rng.integers(0, 2**64, 200, dtype=np.uint64):
            d = orc.decompose(int(x), base_log, level)
            assert all(-(1 << (base_log - 1)) <= v <= (1 << (base_log - 1)) for v in d)
            rec = sum(v << (64 - base_log * (l + 1)) for l, v in enumerate(d)) % 2**64
            err = (rec - int(x)) % 2**64
            err = min(err, 2**64 - err)
            assert err <= 1 << (64 - base_log * level - 1)


def test_modswitch(orc):
    assert orc.modswitch(0, 2048) == 0
    assert orc.modswitch(2**63, 2048) == 2048
    assert orc.modswitch(2**64 - 1, 2048) == 0          # rounds up and wraps
    assert orc.modswitch(2**52, 2048) == 1
    assert orc.modswitch(2**51 - 1, 2048) == 0


def test_fft_against_exact(orc, rng):
    a = rng.integers(0, 2**64, 2048, dtype=np.uint64)
    b = rng.integers(-2**22, 2**22, 2048, dtype=np.int64)
    d = (orc.negacyclic_mul_fft(a, b) - orc.negacyclic_mul_exact(a, b)).astype(np.int64)
    assert np.abs(d).max() < 2**43       # f64 FFT error floor for 2^64 x 2^22 x 2048 terms


def test_encrypt_decrypt_roundtrip(orc, oracle_keys, rng):
    K = oracle_keys("toy")
    m = rng.integers(0, 32, 100).astype(np.uint64)          # includes the padding bit
    assert (K.decrypt_msgs(K.encrypt_msgs(m)) == m).all()


def test_keyswitch_preserves_message(orc, oracle_keys, rng):
    K = oracle_keys("toy")
    m = rng.integers(0, 16, 64).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m))
    assert (K.decode(K.phase_small(small)) == m).all()


@pytest.mark.parametrize("preset", ["toy"])
def test_pbs_evaluates_luts(orc, oracle_keys, rng, preset):
    K = oracle_keys(preset)
    tables = np.stack([np.arange(16), (np.arange(16) ** 2) % 16, rng.integers(0, 16, 16)]).astype(np.uint64)
    luts = np.stack([K.make_lut(t) for t in tables])
    m = np.tile(np.arange(16), 3).astype(np.uint64)
    idx = np.repeat(np.arange(3), 16).astype(np.uint32)
    out = K.ks_pbs(K.encrypt_msgs(m), luts, idx)
    assert (K.decrypt_msgs(out) == tables[idx, m]).all()


def test_pbs_full_params_noise_within_budget(orc, oracle_keys, rng):
    """PARAM_MESSAGE_2_CARRY_2 (n = 834): output noise well inside the decoding radius 2^-5."""
    K = oracle_keys("2_2_gaussian")
    m = rng.integers(0, 16, 16).astype(np.uint64)
    out = K.ks_pbs(K.encrypt_msgs(m), K.make_lut(np.arange(16)))
    assert (K.decrypt_msgs(out) == m).all()
    err = (K.phase_big(out) - K.encode(m)).astype(np.int64).astype(np.float64) / 2.0**64
    assert err.std() < 2.0**-13


def test_bivariate_lut_packing(orc, oracle_keys, rng):
    """lhs*4 + rhs packing used by every bivariate radix op (mul lsb/msb, carry logic)."""
    K = oracle_keys("toy")
    lhs = rng.integers(0, 4, 32).astype(np.uint64)
    rhs = rng.integers(0, 4, 32).astype(np.uint64)
    a, b = K.encrypt_msgs(lhs, stream=0), K.encrypt_msgs(rhs, stream=1000)
    packed = (a * np.uint64(4) + b).astype(np.uint64)
    lsb = K.make_lut([((i >> 2) * (i & 3)) % 4 for i in range(16)])
    msb = K.make_lut([((i >> 2) * (i & 3)) // 4 for i in range(16)])
    assert (K.decrypt_msgs(K.ks_pbs(packed, lsb)) == (lhs * rhs) % 4).all()
    assert (K.decrypt_msgs(K.ks_pbs(packed, msb)) == (lhs * rhs) // 4).all()
