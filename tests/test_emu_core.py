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
    E.emu_negacyclic_mul.argtypes = [vp, vp, vp]
    E.emu_convert_bsk.argtypes = [C.c_int, vp, vp]
    E.emu_blind_rotate.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    E.emu_init()
    return E


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_emu_fft_matches_exact_product(emu, orc, rng):
    a = rng.integers(0, 2**64, 2048, dtype=np.uint64)
    b = rng.integers(-2**22, 2**22, 2048, dtype=np.int64)
    c = np.empty_like(a)
    emu.emu_negacyclic_mul(P(a), P(b), P(c))
    d = (c - orc.negacyclic_mul_exact(a, b)).astype(np.int64)
    assert np.abs(d).max() < 2**43


@pytest.mark.parametrize("acc_bits", [64, 32])
def test_emu_blind_rotation_decrypts_like_oracle(emu, orc, oracle_keys, rng, acc_bits):
    K = oracle_keys("toy")
    n = K.params.lwe_dim
    bf = np.empty(n * 32 * 4 * 32 * 2, dtype=np.float64)
    emu.emu_convert_bsk(n, P(K.bsk), P(bf))
    table = rng.integers(0, 16, 16).astype(np.uint64)
    lut = K.make_lut(table)
    m = rng.integers(0, 16, 48).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m))
    out = np.empty((m.size, 2049), dtype=np.uint64)
    emu.emu_blind_rotate(acc_bits, n, K.params.pbs_base_log, P(bf), P(small), m.size, P(lut), P(out))
    ref = K.pbs(small, lut)
    assert (K.decrypt_msgs(out) == table[m]).all()
    assert (K.decrypt_msgs(ref) == table[m]).all()
    e_emu = (K.phase_big(out) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    e_ref = (K.phase_big(ref) - K.encode(table[m])).astype(np.int64).astype(np.float64)
    assert e_emu.std() < 1.5 * e_ref.std()
