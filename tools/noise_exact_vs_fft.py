"""Where the blind rotation's output noise comes from - measured on the CPU, no GPU needed (toy LWE dimension n = 48,
same GLWE side as the 2_2 parameter set, so per-step noise is that of the real parameters).

Three implementations of the SAME blind rotation on the SAME ciphertexts and keys:
  exact   : every negacyclic product in exact integer arithmetic (orc_negacyclic_mul_exact), round-to-nearest digits
  oracle  : oracle/tfhe_oracle.c (f64 FFT, key transformed by the same FFT)
  kernel  : the CUDA kernels' arithmetic run lane by lane on the CPU (tests/emu/pbs_emu.cpp), with the Fourier key either
            from the kernels' own f64 FFT or correctly rounded (the long-double statement of csrc/bsk_exact.cu)

Result (2 048 / 4 096 samples, profiles/r02_noise_exact_vs_fft.log): exact 0.81 x oracle; kernel with FFT key 1.07-1.11 x
oracle; kernel with the correctly rounded key 1.00 x oracle.  Reading: with l_pbs = 1 the floating-point error of the
Fourier-domain product is a fifth to a quarter of the output variance, because the error in the accumulator's mask is
amplified by the secret key (x (1 + N/2)); a third of it belongs to the key's spectrum and disappears when the key is
transformed exactly, once, at upload.

    python tools/noise_exact_vs_fft.py [samples]      (8 processes, about 5 minutes per 2 048 exact samples)
"""
import ctypes as C
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

N = 2048
K = bsk = lut = table = None


def rot(p, t):
    t %= 2 * N
    idx = (np.arange(N) + 2 * N - t) % (2 * N)
    with np.errstate(over="ignore"):
        return np.where(idx < N, p[idx % N], (np.uint64(0) - p[idx % N]).astype(np.uint64))


def init():
    global K, bsk, lut, table
    K = orc.Keys(orc.preset("toy"), 1)
    bsk = K.bsk.reshape(K.params.lwe_dim, 2, 1, 2, N)
    table = (np.arange(16) * 7 + 3) % 16
    lut = K.make_lut(table)


def exact_blind_rotate(ct):
    """acc <- acc + GGSW_i (x) (X^a acc - acc), every product exact; then sample extraction of coefficient 0"""
    n = K.params.lwe_dim
    bt = orc.modswitch(int(ct[n]), N)
    acc = [np.zeros(N, dtype=np.uint64), rot(lut, (2 * N - bt) % (2 * N))]
    for i in range(n):
        at = orc.modswitch(int(ct[i]), N)
        if not at:
            continue
        with np.errstate(over="ignore"):
            diff = [(rot(acc[q], at) - acc[q]).astype(np.uint64) for q in range(2)]
            dec = [((d + np.uint64(1 << 40)).astype(np.uint64).view(np.int64) >> 41).astype(np.int64) for d in diff]
            new = [acc[q].copy() for q in range(2)]
            for q in range(2):
                for p in range(2):
                    new[q] = (new[q] + orc.negacyclic_mul_exact(bsk[i, p, 0, q], dec[p])).astype(np.uint64)
        acc = new
    ex = np.empty(2049, dtype=np.uint64)
    ex[0] = acc[0][0]
    with np.errstate(over="ignore"):
        ex[1:N] = (np.uint64(0) - acc[0][N - 1:0:-1]).astype(np.uint64)
    ex[N] = acc[1][0]
    return ex


def exact_fourier_key(n):
    """long-double statement of csrc/bsk_exact.cu in the ring kernel's layout [n][slot r][g][lane]: k = lane + 32 brev5(r)"""
    brev5 = lambda v: ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4)
    ld = np.longdouble
    pi = ld("3.14159265358979323846264338327950288")
    e = np.arange(4096).astype(ld)
    cosv, sinv = np.cos(2 * pi * e / 4096), np.sin(2 * pi * e / 4096)
    expo = np.outer(np.arange(1024), 4 * np.arange(1024) + 1) % 4096
    Cm, Sm = cosv[expo], sinv[expo]
    polys = K.bsk.reshape(n * 4, N).view(np.int64)
    out = np.empty((n * 4, 32, 32, 2), dtype=np.float64)
    for pidx in range(n * 4):
        zr, zi = polys[pidx, :1024].astype(ld), polys[pidx, 1024:].astype(ld)
        Xr, Xi = zr @ Cm - zi @ Sm, zr @ Sm + zi @ Cm
        for r in range(32):
            ks = np.arange(32) + 32 * brev5(r)
            out[pidx, r, :, 0], out[pidx, r, :, 1] = Xr[ks].astype(np.float64), Xi[ks].astype(np.float64)
    return out.reshape(n, 4, 32, 32, 2).transpose(0, 2, 1, 3, 4).copy().reshape(-1)


def main():
    cnt = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    init()
    n = K.params.lwe_dim
    rng = np.random.default_rng(3)
    m = rng.integers(0, 16, cnt).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m, seed=5))
    exp = table[m].astype(np.uint64)
    noise = lambda out: (K.phase_big(out) - K.encode(exp)).astype(np.int64).astype(np.float64) / 2.0**64
    t0 = time.time()
    with mp.get_context("spawn").Pool(8, initializer=init) as pool:      # spawn: the oracle's OpenMP runtime does not survive fork
        outs = pool.map(exact_blind_rotate, list(small), chunksize=8)
    v_exact = noise(np.stack(outs)).var()
    v_orc = noise(K.pbs(small, lut)).var()
    print("%d samples, exact arithmetic %.0f s" % (cnt, time.time() - t0))
    print("oracle (f64 FFT)            var 2^%.3f" % np.log2(v_orc))
    print("exact integer products      var 2^%.3f   ratio to oracle %.3f" % (np.log2(v_exact), v_exact / v_orc))
    subprocess_emu = os.path.join(ROOT, "tests", "emu", "libpbs_emu.so")
    E = C.CDLL(subprocess_emu)
    vp = C.c_void_p
    E.emu_convert_bsk.argtypes = [C.c_int, vp, vp]
    E.emu_blind_rotate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    E.emu_init()
    P = lambda a: a.ctypes.data_as(vp)
    key_fft = np.empty(n * 32 * 4 * 32 * 2, dtype=np.float64)
    E.emu_convert_bsk(n, P(K.bsk), P(key_fft))
    key_exact = exact_fourier_key(n)
    for name, key in (("kernel arithmetic, FFT key  ", key_fft), ("kernel arithmetic, exact key", key_exact)):
        for acc, form in ((32, 1), (64, 0)):
            o = np.empty((cnt, 2049), dtype=np.uint64)
            E.emu_blind_rotate(acc, form, n, 23, P(key), P(small), cnt, P(lut), P(o))
            print("%s acc %d: ratio to oracle %.3f" % (name, acc, noise(o).var() / v_orc))


if __name__ == "__main__":
    main()
