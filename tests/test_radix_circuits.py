"""Every radix circuit of fhe_sign_b200/csrc/radix.cpp on the CPU mock backend: results against
Python integers (the plaintext contract of the reference's FheUint operators), no
bootstrap input ever touching the padding bit, noise budget respected (the evaluator throws)."""
import random

import pytest

from mock_radix import MockRadix


@pytest.fixture(scope="module")
def M():
    return MockRadix()


def _rand(rnd, bits):
    k = rnd.choice([0, 1, 2, 3])
    if k == 0:
        return rnd.getrandbits(bits)
    if k == 1:
        return (1 << bits) - 1 - rnd.getrandbits(bits // 4)
    if k == 2:
        return rnd.getrandbits(bits // 2)
    return rnd.choice([0, 1, (1 << bits) - 1, 1 << (bits - 1)])


@pytest.mark.parametrize("bits", [8, 32, 64])
def test_add_sub_mul_wrap(M, bits):
    rnd = random.Random(bits)
    n, mask = bits // 2, (1 << bits) - 1
    for _ in range(40):
        x, y = _rand(rnd, bits), _rand(rnd, bits)
        a, b = M.enc(x, n), M.enc(y, n)
        assert M.dec(a + b) == (x + y) & mask
        assert M.dec(a - b) == (x - y) & mask
        assert M.dec(a * b) == (x * y) & mask
    assert M.counters()["violations"] == 0


def test_reference_kats_on_u32_and_u64(M):
    """decrypted known answers of the reference's tests (src/biguint.rs:429-527)."""
    F = 0xFFFFFFFF
    a64 = lambda v: M.api.cast(M.enc(v, 16), 32)
    s = a64(5) + a64(3)
    assert M.dec(M.api.cast(s >> 32, 16)) == 0 and M.dec(M.api.cast(s & F, 16)) == 8
    s = a64(F) + a64(1)
    assert M.dec(M.api.cast(s >> 32, 16)) == 1 and M.dec(M.api.cast(s & F, 16)) == 0
    s = a64(F) + a64(F)
    assert M.dec(M.api.cast(s >> 32, 16)) == 1 and M.dec(M.api.cast(s & F, 16)) == 0xFFFFFFFE
    # FheUint32 overflow wraps; a shift by the full width is a shift by 0 (src/biguint.rs:469-499)
    w = M.enc(F, 16) + M.enc(1, 16)
    assert M.dec(w) == 0
    t = M.enc(F, 16) + M.enc(F, 16)
    assert M.dec(t >> 32) == 4294967294 and M.dec(t & F) == 4294967294
    p = a64(F) * a64(2)
    assert M.dec(p >> 32) == 1 and M.dec(p & F) == 0xFFFFFFFE
    # scalar ops of src/schnorr.rs:575-607
    assert M.dec(M.enc(123, 16) * 456) == 56088 and M.dec(M.enc(123, 16) + 456) == 579
    assert M.counters()["violations"] == 0


def test_perf_test_op_chain(M):
    """src/perf_test.rs:14-75 with its operands: ((1344 >> 5) as u8).min(7) & 1 == 1, 1344 / 5 == 268."""
    a, b, c = M.enc(1344, 16), M.enc(5, 16), M.enc(7, 4)
    assert M.dec(a + b) == 1349
    assert M.dec(a * b) == 6720
    sh = a >> b
    assert M.dec(sh) == 42
    as8 = M.api.cast(sh, 4)
    mn = M.api.min(as8, c)
    assert M.dec(mn) == 7
    assert M.dec(mn & 1) == 1
    assert M.dec(a // 5) == 268
    assert M.counters()["violations"] == 0


@pytest.mark.parametrize("bits", [8, 32])
def test_shifts(M, bits):
    rnd = random.Random(7)
    n, mask = bits // 2, (1 << bits) - 1
    for _ in range(12):
        x = _rand(rnd, bits)
        a = M.enc(x, n)
        for s in (0, 1, 2, 3, bits // 2 - 1, bits - 1, bits, bits + 3):
            assert M.dec(a >> s) == x >> (s % bits), (x, s)
            assert M.dec(a << s) == (x << (s % bits)) & mask, (x, s)
        amt = rnd.getrandbits(bits)
        e = M.enc(amt, n)
        assert M.dec(a >> e) == x >> (amt % bits)
        assert M.dec(a << e) == (x << (amt % bits)) & mask
    assert M.counters()["violations"] == 0


def test_compare_select_minmax(M):
    rnd = random.Random(3)
    for bits in (8, 32):
        n = bits // 2
        for _ in range(30):
            x, y = _rand(rnd, bits), _rand(rnd, bits)
            if rnd.random() < 0.2:
                y = x
            a, b = M.enc(x, n), M.enc(y, n)
            assert M.dec(M.api.lt(a, b)) == int(x < y)
            assert M.dec(M.api.eq(a, b)) == int(x == y)
            assert M.dec(M.api.min(a, b)) == min(x, y)
            assert M.dec(M.api.max(a, b)) == max(x, y)
            assert M.dec(M.api.select(M.api.lt(a, b), a, b)) == min(x, y)
    assert M.counters()["violations"] == 0


def test_bitops_and_masks(M):
    rnd = random.Random(4)
    for _ in range(20):
        x, y = rnd.getrandbits(32), rnd.getrandbits(32)
        a, b = M.enc(x, 16), M.enc(y, 16)
        assert M.dec(a & b) == x & y
        assert M.dec(M.api.binary("or_", a, b)) == x | y
        assert M.dec(M.api.binary("xor", a, b)) == x ^ y
        assert M.dec(a & y) == x & y
    assert M.counters()["violations"] == 0


@pytest.mark.parametrize("bits", [8, 32, 64])
def test_scalar_mul_div_rem(M, bits):
    rnd = random.Random(bits + 1)
    n, mask = bits // 2, (1 << bits) - 1
    for _ in range(25):
        x = _rand(rnd, bits)
        a = M.enc(x, n)
        c = _rand(rnd, bits)
        assert M.dec(a * c) == (x * c) & mask
        assert M.dec(a + c) == (x + c) & mask
        d = rnd.choice([1, 2, 3, 5, 7, 10, 255, 256, 641, (1 << (bits - 1)) + 1, mask, rnd.getrandbits(bits) | 1, rnd.getrandbits(bits // 2) + 1])
        assert M.dec(a // d) == x // d, (x, d)
        assert M.dec(a % d) == x % d, (x, d)
    with pytest.raises(RuntimeError):
        M.enc(5, n) // 0
    assert M.counters()["violations"] == 0


def test_wide_mul_sum_and_256_bit(M):
    rnd = random.Random(9)
    for _ in range(4):
        x, y, z = rnd.getrandbits(256), rnd.getrandbits(256), rnd.getrandbits(256)
        a, b, c = M.enc(x, 128), M.enc(y, 128), M.enc(z, 128)
        prod = M.api.mul_wide(a, b, 256)
        assert M.dec(prod) == x * y
        s = M.api.sum([prod, M.api.cast(c, 257)], 257)
        assert M.dec(s) == x * y + z
    ops = [M.enc(rnd.getrandbits(64), 32) for _ in range(13)]
    vals = [M.dec(o) for o in ops]
    assert M.dec(M.api.sum(ops, 40)) == sum(vals)
    assert M.dec(M.api.sum(ops, 32)) == sum(vals) & (2**64 - 1)
    n_order = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141      # src/scalar.rs:8
    x = rnd.getrandbits(512)
    assert M.dec(M.enc(x, 257) % n_order) == x % n_order
    assert M.counters()["violations"] == 0


def test_slots_are_released(M):
    import gc
    gc.collect()
    before = M.counters()["live_slots"]
    a, b = M.enc(123456789, 16), M.enc(987654321, 16)
    c = a * b + a
    assert M.dec(c) == (123456789 * 987654321 + 123456789) & 0xFFFFFFFF
    del a, b, c
    gc.collect()
    assert M.counters()["live_slots"] == before


def test_trivial_operands_cost_nothing(M):
    p0, _ = M.api.stats()
    a = M.api.trivial(1234, 16)
    b = M.api.trivial(77, 16)
    assert M.dec(a + b) == 1311 and M.dec(a * b) == 1234 * 77
    p1, _ = M.api.stats()
    assert p1 == p0


@pytest.mark.parametrize("bits,levels", [(8, 3), (32, 5), (64, 6), (256, 9)])
def test_carry_scan_depth_and_long_carry_chains(M, bits, levels):
    """The carry scan is radix 3 (propagate_radix3: binary-adder encoding, one lookup joins three segments) while a level
    of twice the width still runs at one ciphertext per SM — 4-, 16- and 32-block operators on one GPU — and radix 2
    above (128 blocks: 9 levels).  Longest possible carry chains in both directions, through add and sub."""
    rnd = random.Random(1000 + bits)
    n, mask = bits // 2, (1 << bits) - 1
    cases = [(mask, 1), (mask, mask), (1, mask), (mask - 1, 1), (mask >> 1, 1), (0, 0), (mask ^ (1 << (bits // 2)), 1 << (bits // 2))]
    cases += [(_rand(rnd, bits), _rand(rnd, bits)) for _ in range(12)]
    for x, y in cases:
        a, b = M.enc(x, n), M.enc(y, n)
        _, l0 = M.api.stats()
        s = a + b
        _, l1 = M.api.stats()
        assert M.dec(s) == (x + y) & mask
        assert l1 - l0 <= levels
        assert M.dec(a - b) == (x - y) & mask
    assert M.counters()["violations"] == 0


def test_radix3_scan_on_wide_integers_as_an_eight_gpu_node_schedules_them():
    """With level sharding over 8 ranks the radix-3 scan covers everything up to 592 blocks.  Single process: a
    world-8 exchange whose threshold is never reached (nothing is actually cut), so only the schedule changes:
    128 blocks in 2 + 5 levels, 512 blocks in 2 + 6, longest carry chains included."""
    import ctypes as C
    from fhe_sign_b200.distributed import EXCHANGE_FN
    M8 = MockRadix()
    cb = EXCHANGE_FN(lambda user, buf, nbytes: 0)
    dummy = (C.c_uint64 * 4)()
    assert M8.L.fsc_set_level_exchange(M8.ctx, 0, 8, 1 << 30, C.cast(dummy, C.c_void_p), 32, cb, None) == 0
    rnd = random.Random(88)
    for bits, levels in ((256, 7), (1024, 8)):
        n, mask = bits // 2, (1 << bits) - 1
        cases = [(mask, 1), (mask, mask), (mask >> 1, 1), (mask ^ (1 << (bits // 2)), 1 << (bits // 2))]
        cases += [(_rand(rnd, bits), _rand(rnd, bits)) for _ in range(6)]
        for x, y in cases:
            a, b = M8.enc(x, n), M8.enc(y, n)
            _, l0 = M8.api.stats()
            s = a + b
            _, l1 = M8.api.stats()
            assert M8.dec(s) == (x + y) & mask
            assert l1 - l0 <= levels, (bits, l1 - l0)
            assert M8.dec(a - b) == (x - y) & mask
    x, y = _rand(rnd, 256), _rand(rnd, 256)
    assert M8.dec(M8.api.mul_wide(M8.enc(x, 128), M8.enc(y, 128), 256)) == x * y
    assert M8.counters()["violations"] == 0
    assert M8.api.sharded_levels() == 0


def test_comparison_tree_all_sizes_and_depth(M):
    """order_code's radix-3 tree (binary-adder encoding of "less / equal / greater", every node emitted once in the form
    its next role needs, lone nodes passed through): every width from 1 to 40 blocks and a few wide ones, equal and
    one-bit-apart operands included; min / max select on the code directly (u8 min: 4 levels, 128 blocks: 7)."""
    rnd = random.Random(4242)
    for nb in list(range(1, 41)) + [64, 100, 128, 257]:
        bits = 2 * nb
        for t in range(6):
            x = rnd.getrandbits(bits)
            y = x if t % 3 == 0 else (x ^ (1 << rnd.randrange(bits)) if t % 3 == 1 else rnd.getrandbits(bits))
            a, b = M.enc(x, nb), M.enc(y, nb)
            assert M.dec(M.api.lt(a, b)) == (1 if x < y else 0), (nb, x, y)
            _, l0 = M.api.stats()
            mn = M.api.min(a, b)
            _, l1 = M.api.stats()
            assert M.dec(mn) == min(x, y) and M.dec(M.api.max(a, b)) == max(x, y), (nb, x, y)
            if nb == 4:
                assert l1 - l0 <= 4
            if nb == 128:
                assert l1 - l0 <= 7
    assert M.counters()["violations"] == 0


def test_rem_by_pseudo_mersenne_moduli_folds(M):
    """scalar_rem folds hi 2^k + lo -> hi c + lo for moduli 2^k - c with a small c (secp256k1's group order n and field
    prime p, 2^64 - 59, ...): edge values around multiples of the modulus, and the cost of the reference's missing
    `mod n` step (514-bit value, scalar.rs:8) stays far below the general quotient-and-multiply-back path (75 k PBS)."""
    rnd = random.Random(606)
    n_order = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
    p_field = 2**256 - 2**32 - 977
    for d, nb in ((n_order, 257), (n_order, 129), (p_field, 256), (2**64 - 59, 64), (2**32 - 5, 32), (2**16 - 15, 16)):
        bits = 2 * nb
        vals = [(1 << bits) - 1, d, d - 1, d + 1, 2 * d - 1, 2 * d, 0, d * ((1 << bits) // d), d * ((1 << bits) // d) - 1]
        vals += [rnd.getrandbits(bits) for _ in range(3)]
        for v in vals:
            v %= 1 << bits
            p0, _ = M.api.stats()
            r = M.enc(v, nb) % d
            p1, _ = M.api.stats()
            assert M.dec(r) == v % d, (hex(d), hex(v))
            if nb == 257:
                assert p1 - p0 < 20000
    assert M.counters()["violations"] == 0


def test_variance_unit_noise_bookkeeping_is_a_relaxation():
    """FSC_RADIX_NOISE=variance (the default since the GPU noise measurement of round 2) against =linear: same decrypted results, fewer bootstraps in a wide product
    (column-sum chunks may hold more low-degree terms).  The switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, random; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from mock_radix import MockRadix\n"
        "m = MockRadix(); rnd = random.Random(5)\n"
        "x, y, z = rnd.getrandbits(128), rnd.getrandbits(128), rnd.getrandbits(128)\n"
        "p0, _ = m.api.stats()\n"
        "s = m.api.mul_add_wide(m.enc(x, 64), m.enc(y, 64), m.enc(z, 64), 130)\n"
        "p1, _ = m.api.stats()\n"
        "assert m.dec(s) == x * y + z and m.counters()['violations'] == 0\n"
        "a = m.enc(x & 0xFFFFFFFF, 16)\n"
        "assert m.dec(a // 5) == (x & 0xFFFFFFFF) // 5 and m.dec(a * a) == ((x & 0xFFFFFFFF) ** 2) & 0xFFFFFFFF\n"
        "print(p1 - p0)\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    counts = {}
    for mode in ("linear", "variance"):
        env = dict(os.environ, FSC_RADIX_NOISE=mode)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        counts[mode] = int(out.stdout.strip().splitlines()[-1])
    assert counts["variance"] < counts["linear"]
