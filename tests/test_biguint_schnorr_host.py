"""BigUintFHE / sign_fhe_with_k0 host mirrors on the CPU mock backend: the reference's own known
answers (src/biguint.rs:274-527, src/schnorr.rs:440-492), the digit-level replay of biguint.rs as the
oracle (oracle/schnorr_model.py), and the golden BIP-340 signing vectors (tests/golden/)."""
import json
import os
import random

import pytest

from fhe_sign_b200 import biguint as bg
from fhe_sign_b200 import schnorr
from fhe_sign_b200.biguint import BigUintFHE
from mock_radix import MockClientKey, MockRadix
from oracle import schnorr_model as sm

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "schnorr_vectors.json")))
F = 0xFFFFFFFF


@pytest.fixture(scope="module")
def ck():
    m = MockRadix()
    bg.set_server_key(m)
    return MockClientKey(m)


def digits(b, ck):
    return [ck.decrypt(d, None) for d in b.digits]


def test_reference_biguint_kats(ck):
    two, three = BigUintFHE.from_u32(2, ck), BigUintFHE.from_u32(3, ck)
    assert (two * three).to_biguint(ck) == 6                                         # :274-292
    assert BigUintFHE.new(123456789123456789, ck).to_biguint(ck) == 123456789123456789    # :295-305
    s = BigUintFHE.from_u32(F, ck) + BigUintFHE.from_u32(1, ck)                      # :308-333
    assert digits(s, ck) == [0, 1]
    p = BigUintFHE.from_u32(F, ck) * BigUintFHE.from_u32(2, ck)                      # :336-351
    assert digits(p, ck) == [0xFFFFFFFE, 1]
    p = BigUintFHE.from_u32(F, ck) * BigUintFHE.from_u32(F, ck)                      # :354-369, :389-404
    assert digits(p, ck) == [1, 0xFFFFFFFE]
    s = BigUintFHE.from_u32(F, ck) + BigUintFHE.from_u32(F, ck)                      # :372-386
    assert digits(s, ck) == [0xFFFFFFFE, 1]
    a, b = 123456789123456789, 987654321987654321                                    # :407-426
    A, B = BigUintFHE.new(a, ck), BigUintFHE.new(b, ck)
    assert (A.clone() + B.clone()).to_biguint(ck) == a + b
    assert (A * B).to_biguint(ck) == a * b
    assert BigUintFHE.zero(ck).to_biguint(ck) == 0 and BigUintFHE.one(ck).decrypt_to_u32(ck) == 1
    assert BigUintFHE.new(a, ck).decrypt_to_u64(ck) == a and BigUintFHE.new(a, ck).decrypt_to_u32(ck) is None
    assert (BigUintFHE.zero(ck) * A).digits == [] and digits(BigUintFHE.zero(ck) + A, ck) == sm.to_u32_digits(a)


def test_faithful_schedule_matches_digit_replay(ck):
    """including inputs where the reference's Mul drops a carry (src/biguint.rs:247-249)."""
    rnd = random.Random(5)
    cases = [([F] * 3, [F] * 3), ([F, F], [F, F, F])]
    for _ in range(6):
        la, lb = rnd.randint(1, 3), rnd.randint(1, 3)
        cases.append(([rnd.choice([F, rnd.getrandbits(32)]) for _ in range(la)], [rnd.choice([F, rnd.getrandbits(32)]) for _ in range(lb)]))
    dropped = 0
    for da, db in cases:
        a, b = sm.from_digits(da), sm.from_digits(db)
        A, B = BigUintFHE.new(a, ck), BigUintFHE.new(b, ck)
        want_mul = sm.biguint_mul(sm.to_u32_digits(a), sm.to_u32_digits(b))
        assert digits(A.clone() * B.clone(), ck) == want_mul
        dropped += sm.from_digits(want_mul) != a * b
        assert digits(A + B, ck) == sm.biguint_add(sm.to_u32_digits(a), sm.to_u32_digits(b))
    assert dropped >= 1          # the adversarial cases really exercise the dropped-carry path


def test_fused_schedule_is_the_true_sum(ck):
    rnd = random.Random(6)
    for _ in range(3):
        k, e, d = rnd.getrandbits(256), rnd.getrandbits(256), rnd.getrandbits(rnd.choice([32, 256]))
        r = BigUintFHE.mul_add_fused(BigUintFHE.new(k, ck), BigUintFHE.new(e, ck), BigUintFHE.new(d, ck))
        assert r.to_biguint(ck) == k + e * d
        assert len(r.digits) == max(len(sm.to_u32_digits(k)), len(sm.to_u32_digits(e)) + len(sm.to_u32_digits(d))) + 1


def test_golden_vectors_are_pinned_by_the_model():
    """the committed fixtures equal a fresh replay and the reference's own CSV where the reference is right."""
    assert [v["index"] for v in GOLDEN] == [0, 1, 2, 3, 15, 16, 17, 18]
    for v in GOLDEN:
        d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
        assert sm.sign_with_k0(msg, k0, d).hex().upper() == v["reference_signature"]
        if v["reference_matches_csv"]:
            assert v["reference_signature"] == v["csv_signature"].upper()
            assert sm.verify(msg, int(v["public_key"], 16), bytes.fromhex(v["csv_signature"]))
    assert [v["index"] for v in GOLDEN if not v["reference_matches_csv"]] == [3]


@pytest.mark.parametrize("fused", [False, True])
def test_sign_fhe_with_k0_all_vectors(ck, fused):
    """src/schnorr.rs:469-492 generalised to every signing row: sign_fhe_with_k0 == sign_with_k0."""
    for v in GOLDEN:
        if not fused and v["index"] not in (0, 1):       # faithful 8x8 is 64 sequential iterations; two rows suffice on the mock
            continue
        d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
        sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck, fused=fused)
        assert sig.to_bytes().hex().upper() == v["reference_signature"]
        assert sig.to_bytes() == sm.sign_with_k0(msg, k0, d)


def test_sign_fhe_known_answer(ck):
    """src/schnorr.rs:440-466 (test_schnorr_fhe): sign_fhe on BIP-340 vector 0 with aux_rand = 0 gives the published signature,
    through the faithful and through the fused schedule."""
    expected = "E907831F80848D1069A5371B402410364BDF1C5F8307B0084C55F1CE2DCA821525F66A4A85EA8B71E482A74F382D2CE5EBEEE8FDB2172F477DF4900D310536C0"
    for fused in (False, True):
        sig = schnorr.sign_fhe(bytes(32), bytes(32), 3, ck, fused=fused)
        assert sig.to_bytes().hex().upper() == expected


def test_sign_with_the_public_challenge_variant(ck):
    """k + e*d with the challenge e handed over in plaintext (it is public: verify recomputes it) and d, k encrypted:
    fsc_radix_scalar_mul_add_wide.  Same signature bytes as the reference's plaintext twin on every signing row, and about a third
    of the bootstraps of the ciphertext x ciphertext schedule."""
    api = bg._api()
    for v in GOLDEN:
        d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
        p0, _ = api.stats()
        sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck, public_challenge=True)
        p1, _ = api.stats()
        assert sig.to_bytes().hex().upper() == v["reference_signature"]
        if v["index"] == 1:
            assert p1 - p0 < 20000, p1 - p0


def test_sign_with_the_reduction_under_encryption(ck):
    """`s = (k + e d) mod n` entirely under encryption (folding reduction by the secp256k1 order), then the
    same signature bytes as the reference's plaintext `% n` (src/schnorr.rs:276)."""
    for v in GOLDEN[:3]:
        d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
        sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck, fused=True, reduce_encrypted=True)
        assert sig.to_bytes().hex().upper() == v["reference_signature"]
    x = BigUintFHE.new(sm.N * 5 + 12345, ck) if hasattr(sm, "N") else BigUintFHE.new(schnorr.N * 5 + 12345, ck)
    assert x.rem_scalar(schnorr.N).to_biguint(ck) == 12345
