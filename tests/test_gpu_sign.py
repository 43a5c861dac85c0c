"""End-to-end parity on the GPU: BigUintFHE and Schnorr::sign_fhe_with_k0 against the reference's own
plaintext twin (sign_with_k0) on the BIP-340 signing vectors (tests/golden/schnorr_vectors.json)."""
import json
import os
import time

import pytest

from fhe_sign_b200 import biguint as bg
from fhe_sign_b200 import schnorr
from fhe_sign_b200.biguint import BigUintFHE
from oracle_client import OracleClientKey

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "schnorr_vectors.json")))
F = 0xFFFFFFFF


@pytest.fixture(scope="module", params=[32, 64], ids=["acc32", "acc64"])
def ck(request, gpu_ctx, oracle_keys):
    """the whole signing suite at the product's default accumulator width (32) and at the reference's (64)"""
    client = OracleClientKey(oracle_keys("2_2_gaussian"))
    client.ctx = gpu_ctx("2_2_gaussian", request.param)
    return client


@pytest.fixture(autouse=True)
def _install_server_key(request):
    """tfhe::set_server_key before every test that uses `ck` (another test may have installed its own context)"""
    if "ck" in request.fixturenames:
        bg.set_server_key(request.getfixturevalue("ck").ctx)


def test_biguint_kats_on_gpu(ck):
    """src/biguint.rs:274-426 with real ciphertexts."""
    assert (BigUintFHE.from_u32(2, ck) * BigUintFHE.from_u32(3, ck)).to_biguint(ck) == 6
    s = BigUintFHE.from_u32(F, ck) + BigUintFHE.from_u32(1, ck)
    assert [ck.decrypt(d, bg._api()) for d in s.digits] == [0, 1]
    p = BigUintFHE.from_u32(F, ck) * BigUintFHE.from_u32(F, ck)
    assert [ck.decrypt(d, bg._api()) for d in p.digits] == [1, 0xFFFFFFFE]
    a, b = 123456789123456789, 987654321987654321
    assert (BigUintFHE.new(a, ck) + BigUintFHE.new(b, ck)).to_biguint(ck) == a + b
    assert (BigUintFHE.new(a, ck) * BigUintFHE.new(b, ck)).to_biguint(ck) == a * b


def test_sign_fhe_known_answer_on_gpu(ck):
    """src/schnorr.rs:440-466 (test_schnorr_fhe): `sign_fhe` (nonce from aux_rand, key encrypted inside) on vector 0 gives the
    signature the reference's test pins, on real ciphertexts."""
    expected = "E907831F80848D1069A5371B402410364BDF1C5F8307B0084C55F1CE2DCA821525F66A4A85EA8B71E482A74F382D2CE5EBEEE8FDB2172F477DF4900D310536C0"
    sig = schnorr.sign_fhe(bytes(32), bytes(32), 3, ck, fused=True)
    assert sig.to_bytes().hex().upper() == expected


def test_sign_fhe_with_k0_vector0_faithful(ck):
    """src/schnorr.rs:469-492: vector 0 (d = 3) through the reference's own op-for-op schedule."""
    v = GOLDEN[0]
    d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
    p0, l0 = bg._api().stats()
    t0 = time.perf_counter()
    sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck)
    dt = time.perf_counter() - t0
    p1, l1 = bg._api().stats()
    print("faithful sign, vector 0: %.2f s, %d PBS in %d levels" % (dt, p1 - p0, l1 - l0))
    assert sig.to_bytes().hex().upper() == v["reference_signature"] == v["csv_signature"].upper()


def test_sign_fhe_with_k0_vector1_faithful_8x8_digits(ck, request):
    """One full-width signature (8 x 8 digits: 64 iterations of src/biguint.rs:214-254, then the ripple Add of :123-191)
    through the reference's own op-for-op schedule - every cast, every FheUint64 add / mul / >> 32 / & 0xFFFFFFFF and the
    wrapping u32 add of :247-249 in the reference's order - on real ciphertexts."""
    if "acc64" in request.node.name:
        pytest.skip("the faithful 8 x 8 schedule runs once, at the default accumulator width")
    v = GOLDEN[1]
    d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
    p0, l0 = bg._api().stats()
    t0 = time.perf_counter()
    sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck)
    dt = time.perf_counter() - t0
    p1, l1 = bg._api().stats()
    print("faithful sign, vector 1 (8 x 8 digits): %.2f s, %d PBS in %d levels" % (dt, p1 - p0, l1 - l0))
    assert sig.to_bytes().hex().upper() == v["reference_signature"] == v["csv_signature"].upper()


def test_sign_fhe_with_k0_all_vectors_fused(ck):
    """every BIP-340 row with a secret key: sign_fhe_with_k0 == sign_with_k0 (bit-exact signature bytes)."""
    for v in GOLDEN:
        d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
        p0, l0 = bg._api().stats()
        t0 = time.perf_counter()
        sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck, fused=True)
        dt = time.perf_counter() - t0
        p1, l1 = bg._api().stats()
        print("fused sign, vector %d: %.2f s, %d PBS in %d levels" % (v["index"], dt, p1 - p0, l1 - l0))
        assert sig.to_bytes().hex().upper() == v["reference_signature"]
        if v["reference_matches_csv"]:
            assert sig.to_bytes().hex().upper() == v["csv_signature"].upper()


def test_sign_with_public_challenge_on_gpu(ck):
    """the scalar-challenge variant (e in plaintext, d and k encrypted: fsc_radix_scalar_mul_add_wide) on real ciphertexts:
    the reference's signature bytes for every signing row."""
    for v in GOLDEN:
        d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
        p0, l0 = bg._api().stats()
        t0 = time.perf_counter()
        sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck, public_challenge=True)
        dt = time.perf_counter() - t0
        p1, l1 = bg._api().stats()
        print("public-challenge sign, vector %d: %.2f s, %d PBS in %d levels" % (v["index"], dt, p1 - p0, l1 - l0))
        assert sig.to_bytes().hex().upper() == v["reference_signature"]


def test_product_only_pipeline_no_oracle():
    """keys, encryption, GPU evaluation and decryption all through the product's own C ABI (fsc_client_* +
    fsc_radix_*): the oracle is not involved.  Known answers of src/biguint.rs:407-426 and vector 1 signing."""
    import fhe_sign_b200 as fsb
    from fhe_sign_b200.client import generate_keys
    ck, (bsk, ksk) = generate_keys("2_2_tuniform", seed=2024)
    ctx = fsb.Context(fsb.Params.preset("2_2_tuniform", acc_bits=32))
    ctx.upload_keys(bsk, ksk)
    bg.set_server_key(ctx)
    a, b = 123456789123456789, 987654321987654321
    assert (BigUintFHE.new(a, ck) + BigUintFHE.new(b, ck)).to_biguint(ck) == a + b
    assert (BigUintFHE.new(a, ck) * BigUintFHE.new(b, ck)).to_biguint(ck) == a * b
    v = GOLDEN[1]
    d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
    sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck, fused=True)
    assert sig.to_bytes().hex().upper() == v["reference_signature"] == v["csv_signature"].upper()
    ctx.close()
