"""Radix circuits on REAL ciphertexts without a GPU: the product's radix layer (csrc/radix.cpp) over the CPU oracle as
its device (tests/host/oracle_backend.cpp).  The plaintext mock (test_radix_circuits.py) checks the circuits' logic on
thousands of inputs; this checks, on a few, that the same circuits survive real noise - every level is an actual
keyswitch + PBS - and pins the reference's known answers at the ciphertext level on the CPU.
Toy LWE dimension (n = 48; same GLWE side, same plaintext encoding) keeps it to seconds."""
import random

import pytest

from oracle_client import OracleClientKey
from oracle_radix import OracleRadix

F = 0xFFFFFFFF


@pytest.fixture(scope="module")
def cl(oracle_keys):
    K = oracle_keys("toy")
    dev = OracleRadix(K)

    class Client:
        api = dev.radix
        ck = OracleClientKey(K, seed=31)

        def enc(self, v, n):
            return self.ck.encrypt_blocks(v, n, self.api)

        def dec(self, r):
            return self.ck.decrypt(r, self.api)
    yield Client()
    dev.close()


def test_perf_test_chain_on_the_cpu(cl):
    """src/perf_test.rs:14-75 on its own operands, every operator a sequence of real keyswitch + PBS levels."""
    a, b, c = cl.enc(1344, 16), cl.enc(5, 16), cl.enc(7, 4)
    assert cl.dec(a + b) == 1349
    assert cl.dec(a * b) == 6720
    sh = a >> b
    assert cl.dec(sh) == 42
    mn = cl.api.min(cl.api.cast(sh, 4), c)
    assert cl.dec(mn) == 7
    assert cl.dec(mn & 1) == 1
    assert cl.dec(a // 5) == 268


def test_u64_carry_kats_on_the_cpu(cl):
    """src/biguint.rs:429-466, :502-527."""
    a64 = lambda v: cl.api.cast(cl.enc(v, 16), 32)
    for x, y, hi, lo in ((5, 3, 0, 8), (F, 1, 1, 0), (F, F, 1, 0xFFFFFFFE)):
        s = a64(x) + a64(y)
        assert cl.dec(cl.api.cast(s >> 32, 16)) == hi and cl.dec(cl.api.cast(s & F, 16)) == lo
    p = a64(F) * a64(F)
    assert cl.dec(p) == F * F


def test_biguint_and_fused_k_plus_ed_on_the_cpu(cl):
    """BigUintFHE (host mirror of src/biguint.rs) over the oracle device: the 2-digit known answers of :407-426, and a
    64 x 64-bit k + e*d through the fused schedule against the faithful op-for-op one."""
    from fhe_sign_b200 import biguint as bg
    from fhe_sign_b200.biguint import BigUintFHE

    class Dev:
        radix = cl.api
    bg.set_server_key(Dev)
    ck = cl.ck
    a, b = 123456789123456789, 987654321987654321
    A, B = BigUintFHE.new(a, ck), BigUintFHE.new(b, ck)
    assert (A + B).to_biguint(ck) == a + b
    prod = A * B
    assert prod.to_biguint(ck) == a * b
    k = 0xFEDCBA9876543210
    fused = BigUintFHE.mul_add_fused(BigUintFHE.new(k, ck), A, B)
    assert fused.to_biguint(ck) == k + a * b


def test_random_operators_on_the_cpu(cl):
    rnd = random.Random(5)
    x, y = rnd.getrandbits(32), rnd.getrandbits(32)
    a, b = cl.enc(x, 16), cl.enc(y, 16)
    assert cl.dec(a - b) == (x - y) & F
    assert cl.dec(cl.api.max(a, b)) == max(x, y)
    assert cl.dec(a & b) == x & y
    assert cl.dec(a << b) == (x << (y % 32)) & F
    d = rnd.getrandbits(12) | 1
    assert cl.dec(a % d) == x % d
