"""The warp-resident blind rotation (fhe_sign_b200/csrc/pbs_core.cuh) executed lane by lane on the
CPU by tests/emu/pbs_emu.cpp, checked against the oracle.  This is the no-GPU proof of the kernel's
index, twiddle and rounding logic; the -m gpu tests check the real kernel the same way."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "emu", "libpbs_emu.so")
    src = os.path.join(HERE, "emu", "pbs_emu.cpp")
    core = os.path.join(HERE, "..", "fhe_sign_b200", "csrc", "pbs_core.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(core)) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    E = C.CDLL(so)
    vp = C.c_void_p
    E.emu_negacyclic_mul.argtypes = [C.c_int, vp, vp, vp]
    E.emu_convert_bsk.argtypes = [C.c_int, vp, vp]
    E.emu_blind_rotate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    E.emu_init()
    return E


def P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("form", [0, 1])
def test_emu_fft_matches_exact_product(emu, orc, rng, form):
    """form 0: plain constants (64-bit accumulator ring, pair kernel); form 1: the half-step ring's tangent-form passes."""
    a = rng.integers(0, 2**64, 2048, dtype=np.uint64)
    b = rng.integers(-2**22, 2**22, 2048, dtype=np.int64)
    c = np.empty_like(a)
    emu.emu_negacyclic_mul(form, P(a), P(b), P(c))
    d = (c - orc.negacyclic_mul_exact(a, b)).astype(np.int64)
    assert np.abs(d).max() < 2**43


@pytest.mark.parametrize("acc_bits,form", [(64, 0), (32, 1)])
def test_emu_blind_rotation_decrypts_like_oracle(emu, orc, oracle_keys, rng, acc_bits, form):
    K = oracle_keys("toy")
    n = K.params.lwe_dim
    bf = np.empty(n * 32 * 4 * 32 * 2, dtype=np.float64)
    emu.emu_convert_bsk(n, P(K.bsk), P(bf))
    table = rng.integers(0, 16, 16).astype(np.uint64)
    lut = K.make_lut(table)
    m = rng.integers(0, 16, 48).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m))
    out = np.empty((m.size, 2049), dtype=np.uint64)
    emu.emu_blind_rotate(acc_bits, form, n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out))
    ref = K.pbs(small, lut)
    assert (K.decrypt_msgs(out) == table[m]).all()
    assert (K.decrypt_msgs(ref) == table[m]).all()
    e_emu = (K.phase_big(out) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    e_ref = (K.phase_big(ref) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    assert e_emu.std() < 1.5 * e_ref.std()


# ---- the single-routine "stream" formulation (pbs_core2.cuh / pbs_stream_kernel.cu) ----

@pytest.fixture(scope="module")
def emu2():
    so = os.path.join(HERE, "emu", "libpbs_emu2.so")
    src = os.path.join(HERE, "emu", "pbs_emu2.cpp")
    cores = [os.path.join(HERE, "..", "fhe_sign_b200", "csrc", f) for f in ("pbs_core.cuh", "pbs_core2.cuh")]
    if not os.path.exists(so) or max([os.path.getmtime(src)] + [os.path.getmtime(c) for c in cores]) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    E = C.CDLL(so)
    vp = C.c_void_p
    E.emu2_negacyclic_mul.argtypes = [vp, vp, vp]
    E.emu2_convert_bsk.argtypes = [C.c_int, vp, vp]
    E.emu2_blind_rotate.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    E.emu2_min_cos.restype = C.c_double
    E.emu2_set_uniform.argtypes = [C.c_int]
    return E


def test_stream_frequency_order_is_closed_under_bit_reversal():
    """pbs_core2.cuh freq_at / freq_pos: halves and 4-slot chunks closed under brev5 (the product swaps slots inside a chunk)."""
    def brev5(v):
        return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4)
    seq = [0, 2, 8, 4, 6, 12, 10, 14, 17, 19, 25, 21, 23, 29, 27, 31, 1, 16, 3, 24, 5, 20, 7, 28, 9, 18, 11, 26, 13, 22, 15, 30]
    pos = [0, 16, 1, 18, 3, 20, 4, 22, 2, 24, 6, 26, 5, 28, 7, 30, 17, 8, 25, 9, 21, 11, 29, 12, 19, 10, 27, 14, 23, 13, 31, 15]
    assert sorted(seq) == list(range(32)) and all(pos[k] == seq.index(k) for k in range(32))
    for c in range(8):
        chunk = seq[4 * c:4 * c + 4]
        assert sorted(chunk) == sorted(brev5(x) for x in chunk)


@pytest.mark.parametrize("acc_bits", [64, 32])
def test_stream_uniform_passes_specialised(emu2, orc, oracle_keys, rng, acc_bits):
    """pass32_uniform (compile-time root parameter: trivial constants as additions, tangent level 1) is pass32 on the uniform
    tables up to the rounding of the multiplications it removes: exact products agree, the blind rotation decrypts, same noise."""
    a = rng.integers(0, 2**64, 2048, dtype=np.uint64)
    b = rng.integers(-2**22, 2**22, 2048, dtype=np.int64)
    c = np.empty_like(a)
    K = oracle_keys("toy")
    n = K.params.lwe_dim
    bf = np.empty(n * 32 * 4 * 32 * 2, dtype=np.float64)
    emu2.emu2_convert_bsk(n, P(K.bsk), P(bf))
    table = rng.integers(0, 16, 16).astype(np.uint64)
    lut = K.make_lut(table)
    m = rng.integers(0, 16, 128).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m))
    out, out0 = np.empty((m.size, 2049), dtype=np.uint64), np.empty((m.size, 2049), dtype=np.uint64)
    emu2.emu2_blind_rotate(acc_bits, n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out0))
    emu2.emu2_set_uniform(1)
    try:
        emu2.emu2_negacyclic_mul(P(a), P(b), P(c))
        emu2.emu2_blind_rotate(acc_bits, n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out))
    finally:
        emu2.emu2_set_uniform(0)
    d = (c - orc.negacyclic_mul_exact(a, b)).astype(np.int64)
    assert np.abs(d).max() < 2**43
    assert (K.decrypt_msgs(out) == table[m]).all()
    e1 = (K.phase_big(out) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    e0 = (K.phase_big(out0) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    assert e1.std() < 1.3 * e0.std(), (e1.std(), e0.std())


def test_stream_tangent_constants_are_finite(emu2):
    assert emu2.emu2_min_cos() > 4e-3


def test_stream_fft_matches_exact_product(emu2, orc, rng):
    a = rng.integers(0, 2**64, 2048, dtype=np.uint64)
    b = rng.integers(-2**22, 2**22, 2048, dtype=np.int64)
    c = np.empty_like(a)
    emu2.emu2_negacyclic_mul(P(a), P(b), P(c))
    d = (c - orc.negacyclic_mul_exact(a, b)).astype(np.int64)
    assert np.abs(d).max() < 2**43


@pytest.mark.parametrize("acc_bits", [64, 32])
def test_stream_blind_rotation_decrypts_like_oracle(emu2, orc, oracle_keys, rng, acc_bits):
    K = oracle_keys("toy")
    n = K.params.lwe_dim
    bf = np.empty(n * 32 * 4 * 32 * 2, dtype=np.float64)
    emu2.emu2_convert_bsk(n, P(K.bsk), P(bf))
    table = rng.integers(0, 16, 16).astype(np.uint64)
    lut = K.make_lut(table)
    m = rng.integers(0, 16, 48).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m))
    out = np.empty((m.size, 2049), dtype=np.uint64)
    emu2.emu2_blind_rotate(acc_bits, n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out))
    ref = K.pbs(small, lut)
    assert (K.decrypt_msgs(out) == table[m]).all()
    e_emu = (K.phase_big(out) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    e_ref = (K.phase_big(ref) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    assert e_emu.std() < 1.5 * e_ref.std()


# ---- split formulation: four warps per ciphertext (pbs_core3.cuh, tests/emu/pbs_emu3.cpp) -----------------------
@pytest.fixture(scope="module")
def emu3():
    so = os.path.join(HERE, "emu", "libpbs_emu3.so")
    src = os.path.join(HERE, "emu", "pbs_emu3.cpp")
    cores = [os.path.join(HERE, "..", "fhe_sign_b200", "csrc", f) for f in ("pbs_core.cuh", "pbs_core2.cuh", "pbs_core3.cuh")]
    if not os.path.exists(so) or max(os.path.getmtime(f) for f in [src] + cores) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    E = C.CDLL(so)
    vp = C.c_void_p
    E.emu3_blind_rotate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    return E


@pytest.mark.parametrize("px", [0, 1, 2])
@pytest.mark.parametrize("acc_bits", [64, 32])
def test_split_blind_rotation_matches_stream_formulation(emu2, emu3, orc, oracle_keys, rng, acc_bits, px):
    """The half-pass decomposition is the same arithmetic as pass32 on 32 slots up to the order of the product's
    additions: outputs agree with the stream emulator to rounding noise and decrypt like the oracle.
    px = 1: the Fourier-domain product split between the two warps of a polynomial; px = 2: the same with the half
    index at run time (what the kernel runs)."""
    K = oracle_keys("toy")
    n = K.params.lwe_dim
    bf = np.empty(n * 32 * 4 * 32 * 2, dtype=np.float64)
    emu2.emu2_convert_bsk(n, P(K.bsk), P(bf))
    table = rng.integers(0, 16, 16).astype(np.uint64)
    lut = K.make_lut(table)
    m = rng.integers(0, 16, 32).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m))
    out, out2 = np.empty((m.size, 2049), dtype=np.uint64), np.empty((m.size, 2049), dtype=np.uint64)
    emu3.emu3_blind_rotate(acc_bits, px, n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out))
    emu2.emu2_blind_rotate(acc_bits, n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out2))
    assert (K.decrypt_msgs(out) == table[m]).all()
    e3 = (K.phase_big(out) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    e2 = (K.phase_big(out2) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    assert e3.std() < 1.5 * e2.std()


# ---- quad formulation: the split form for wide batches (pbs_core4.cuh, tests/emu/pbs_emu4.cpp) ---------------------------
@pytest.fixture(scope="module")
def emu4():
    so = os.path.join(HERE, "emu", "libpbs_emu4.so")
    src = os.path.join(HERE, "emu", "pbs_emu4.cpp")
    cores = [os.path.join(HERE, "..", "fhe_sign_b200", "csrc", f) for f in ("pbs_core.cuh", "pbs_core2.cuh", "pbs_core3.cuh", "pbs_core4.cuh")]
    if not os.path.exists(so) or max(os.path.getmtime(f) for f in [src] + cores) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    E = C.CDLL(so)
    vp = C.c_void_p
    E.emu4_blind_rotate.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    return E


def test_quad_transpose_swizzle_is_conflict_free(emu4):
    """[32][32] complex with physical column = column ^ row: no two lanes of a quarter warp share a 16-byte bank group,
    for the row stores and for both kinds of column loads."""
    assert emu4.emu4_transpose_conflicts() == 1


def test_quad_blind_rotation_matches_stream_formulation(emu2, emu4, orc, oracle_keys, rng):
    """Whole level-1 butterflies + the 8-value join, the [parity][index] accumulator columns and the neighbouring-slot
    product are the stream formulation's arithmetic in another order of ownership: same decrypted values, same noise."""
    K = oracle_keys("toy")
    n = K.params.lwe_dim
    bf = np.empty(n * 32 * 4 * 32 * 2, dtype=np.float64)
    emu2.emu2_convert_bsk(n, P(K.bsk), P(bf))
    table = rng.integers(0, 16, 16).astype(np.uint64)
    lut = K.make_lut(table)
    m = rng.integers(0, 16, 32).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m))
    out, out2 = np.empty((m.size, 2049), dtype=np.uint64), np.empty((m.size, 2049), dtype=np.uint64)
    emu4.emu4_blind_rotate(n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out))
    emu2.emu2_blind_rotate(32, n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out2))
    assert (K.decrypt_msgs(out) == table[m]).all()
    e4 = (K.phase_big(out) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    e2 = (K.phase_big(out2) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    assert e4.std() < 1.5 * e2.std()
