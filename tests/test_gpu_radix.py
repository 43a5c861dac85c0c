"""Radix operators on the GPU with real encryptions (keys and client side from the oracle), checked
against the plaintext contract of the reference's operators and its known answers."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PRESET = "2_2_gaussian"


class Client:
    """client side of the reference (FheUintN::try_encrypt / decrypt, src/biguint.rs:26,70), on the oracle"""

    def __init__(self, K, api):
        self.K, self.api, self.stream = K, api, 0

    def enc(self, value, n_blocks):
        digits = np.array([(int(value) >> (2 * i)) & 3 for i in range(n_blocks)], dtype=np.uint64)
        ct = self.K.encrypt_msgs(digits, seed=99, stream=self.stream)
        self.stream += n_blocks
        return self.api.from_lwe(ct)

    def dec(self, r):
        d = self.K.decrypt_msgs(self.api.to_lwe(r))
        assert (d < 4).all(), d
        return sum(int(v) << (2 * i) for i, v in enumerate(d))


@pytest.fixture(scope="module", params=[32, 64], ids=["acc32", "acc64"])
def cl(request, gpu_ctx, oracle_keys):
    """every operator test runs at the product's default accumulator width (32) and at the reference's (64)"""
    ctx = gpu_ctx(PRESET, request.param)
    return Client(oracle_keys(PRESET), ctx.radix)


def test_u32_add_mul_kats(cl):
    F = 0xFFFFFFFF
    a, b = cl.enc(F, 16), cl.enc(1, 16)
    assert cl.dec(a + b) == 0                                    # src/biguint.rs:469-499 wrapping
    assert cl.dec(cl.enc(123, 16) * 456) == 56088                # src/schnorr.rs:575-592
    assert cl.dec(cl.enc(123, 16) + 456) == 579                  # src/schnorr.rs:595-607
    x, y = 123456789, 987654321
    assert cl.dec(cl.enc(x, 16) * cl.enc(y, 16)) == (x * y) & F
    assert cl.dec(cl.enc(x, 16) + cl.enc(y, 16)) == (x + y) & F
    assert cl.dec(cl.enc(x, 16) - cl.enc(y, 16)) == (x - y) & F


def test_u64_carry_extraction_kats(cl):
    """src/biguint.rs:429-466 and :502-527: extract_upper_bits / extract_lower_bits of FheUint64 sums."""
    F = 0xFFFFFFFF
    api = cl.api
    a64 = lambda v: api.cast(cl.enc(v, 16), 32)
    for x, y, hi, lo in ((5, 3, 0, 8), (F, 1, 1, 0), (F, F, 1, 0xFFFFFFFE)):
        s = a64(x) + a64(y)
        assert cl.dec(api.cast(s >> 32, 16)) == hi and cl.dec(api.cast(s & F, 16)) == lo
    p = a64(F) * a64(2)
    assert cl.dec(p >> 32) == 1 and cl.dec(p & F) == 0xFFFFFFFE
    p = a64(F) * a64(F)
    assert cl.dec(p) == F * F


def test_perf_test_chain(cl):
    """src/perf_test.rs:14-75 on its own operands."""
    api = cl.api
    a, b, c = cl.enc(1344, 16), cl.enc(5, 16), cl.enc(7, 4)
    assert cl.dec(a + b) == 1349
    assert cl.dec(a * b) == 6720
    sh = a >> b
    assert cl.dec(sh) == 42
    mn = api.min(api.cast(sh, 4), c)
    assert cl.dec(mn) == 7
    assert cl.dec(mn & 1) == 1
    assert cl.dec(a // 5) == 268


def test_random_ops_u32(cl):
    rnd = random.Random(11)
    F = 0xFFFFFFFF
    for _ in range(3):
        x, y = rnd.getrandbits(32), rnd.getrandbits(32)
        a, b = cl.enc(x, 16), cl.enc(y, 16)
        assert cl.dec(a * b) == (x * y) & F
        assert cl.dec(cl.api.min(a, b)) == min(x, y)
        assert cl.dec(a >> (y % 32)) == x >> (y % 32)
        assert cl.dec(a >> b) == x >> (y % 32)
        assert cl.dec(a & b) == x & y
        d = rnd.getrandbits(20) | 1
        assert cl.dec(a // d) == x // d and cl.dec(a % d) == x % d


def test_256_bit_mul_add(cl):
    """BASELINE configs[2]: one 256 x 256 -> 512-bit product + 256-bit addend (the k + e*d of signing)."""
    rnd = random.Random(12)
    x, y, z = rnd.getrandbits(256), rnd.getrandbits(256), rnd.getrandbits(256)
    a, b, c = cl.enc(x, 128), cl.enc(y, 128), cl.enc(z, 128)
    prod = cl.api.mul_wide(a, b, 256)
    s = cl.api.sum([prod, cl.api.cast(c, 257)], 257)
    assert cl.dec(s) == x * y + z


def test_k_plus_ed_fused_and_reduced_mod_n(cl):
    """src/schnorr.rs:272-276 end to end on ciphertexts: k + e*d as one fsc_radix_mul_add_wide (the addend in the
    product's column sum), then `mod n` under encryption by the folding reduction - the reference
    takes it after decryption.  Also x / 5 and x % (2^32 - 5) through the fix-up-free division and the fold."""
    n_order = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
    rnd = random.Random(13)
    k, e, d = rnd.getrandbits(256) % n_order, rnd.getrandbits(256), rnd.getrandbits(256) % n_order
    s = cl.api.mul_add_wide(cl.enc(e, 128), cl.enc(d, 128), cl.enc(k, 128), 257)
    assert cl.dec(s) == k + e * d
    assert cl.dec(s % n_order) == (k + e * d) % n_order
    x = rnd.getrandbits(32)
    a = cl.enc(x, 16)
    assert cl.dec(a // 5) == x // 5
    assert cl.dec(a % (2**32 - 5)) == x % (2**32 - 5)


def test_256_bit_shift_min_div_at_128_blocks(cl):
    """BASELINE configs[2] / [3] at their named width: the src/perf_test.rs:36,44,54 operators (>> by an encrypted amount,
    min, / 5) on 256-bit values = 128 radix blocks, checked by decryption."""
    rnd = random.Random(14)
    M = 2**256 - 1
    x, y = rnd.getrandbits(256), rnd.getrandbits(256)
    a, b = cl.enc(x, 128), cl.enc(y, 128)
    assert cl.dec(a >> b) == x >> (y % 256)                          # amount taken modulo the width (src/biguint.rs:469-499 rule)
    amt = cl.enc(77, 128)
    assert cl.dec(a >> amt) == x >> 77
    assert cl.dec(a << amt) == (x << 77) & M
    assert cl.dec(cl.api.min(a, b)) == min(x, y)
    assert cl.dec(cl.api.max(a, b)) == max(x, y)
    near = cl.enc(x ^ 1, 128)                                        # operands that differ in the last bit only
    assert cl.dec(cl.api.min(a, near)) == min(x, x ^ 1)
    assert cl.dec(a // 5) == x // 5
    assert cl.dec(a % 5) == x % 5
    assert cl.dec(cl.enc(M, 128) // 5) == M // 5                     # all-ones dividend: every carry chain at full length


def test_256_bit_add_sub_mul_wrapping(cl):
    rnd = random.Random(15)
    M = 2**256
    x, y = rnd.getrandbits(256), rnd.getrandbits(256)
    a, b = cl.enc(x, 128), cl.enc(y, 128)
    assert cl.dec(a + b) == (x + y) % M
    assert cl.dec(a - b) == (x - y) % M
    assert cl.dec(a * b) == (x * y) % M
    assert cl.dec(cl.enc(M - 1, 128) + cl.enc(1, 128)) == 0          # the longest carry chain
