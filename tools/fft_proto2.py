"""numpy statement of the single-routine ("all Cooley-Tukey") negacyclic FFT used by pbs_stream_kernel.cu
(fhe_sign_b200/csrc/pbs_core2.cuh).  One 32-point pass routine serves all four passes of a CMUX step; only its
constant table differs:

  pass(v, g):  out[pos] = sum_j v[j] * (zeta^(g + 128 brev5(pos)))^j      zeta = exp(2 pi i / 4096)

  forward   z_j (j = j1 + 32 j2)  ->  X_k = sum_j z_j zeta^(j (4k+1)),  k = k2 + 32 k1
      pass(g = 32) over j2 at lane j1            -> slot pos holds k2 = brev5(pos)
      transpose                                   -> lane k2, slot j1
      pass(g = 4 k2 + 1) over j1                  -> slot pos holds k1 = brev5(pos)
  inverse   z_j = zeta^(-j) / 1024 * sum_k X_k (zeta^(-4j))^k
      (slot k1 natural, lane k2)
      pass(g = 0) over k1                         -> slot pos holds j1 = -brev5(pos) mod 32
      transpose (row = brev5(pos); lane j1 reads row -j1 mod 32)
      pass(g = -4 j1) over k2                     -> slot pos holds j2 = -brev5(pos) mod 32
      twist by zeta^(-(j1 + 32 j2)) / 1024
"""
import numpy as np

ZN = 4096


def zeta(e):
    return np.exp(2j * np.pi * (np.asarray(e) % ZN) / ZN)


def brev5(v):
    v = np.asarray(v)
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4)


def node_level(ci):
    return 5 if ci >= 8 else 4 if ci >= 4 else 3 if ci >= 2 else 2 if ci == 1 else 1


def node_exponent(ci, g):
    L = node_level(ci)
    t = 0 if L == 1 else ci - (1 << (L - 2))
    return (32 >> L) * g + 64 * int(brev5(2 * t))


def pass32(v, g):
    """v: [lanes, 32]; g: per-lane root parameter (array) -> in-place butterflies lo +- s*hi, odd node = i * even node."""
    g = np.asarray(g)
    v = v.copy()
    for L in range(1, 6):
        half = 16 >> (L - 1)
        for m in range(1 << (L - 1)):
            base = m * 2 * half
            ci = 0 if L == 1 else (1 << (L - 2)) + (m >> 1)
            s = zeta(np.array([node_exponent(ci, int(x)) for x in g]))[:, None]
            if L > 1 and (m & 1):
                s = 1j * s
            lo = v[:, base:base + half].copy()
            hi = v[:, base + half:base + 2 * half] * s
            v[:, base:base + half] = lo + hi
            v[:, base + half:base + 2 * half] = lo - hi
    return v


POS = np.arange(32)
BR = brev5(POS)
NBR = (-BR) % 32
lanes = np.arange(32)


def forward(z):
    v = z.reshape(32, 32).T.copy()                 # [lane j1][slot j2]
    v = pass32(v, np.full(32, 32))                 # slot pos <-> k2 = BR[pos]
    w = np.empty_like(v)
    w[BR[:, None], lanes[None, :]] = v.T           # w[k2][j1]
    v = pass32(w, 4 * lanes + 1)                   # lane k2, slot pos <-> k1 = BR[pos]
    X = np.empty(1024, complex)
    for pos in range(32):
        X[lanes + 32 * BR[pos]] = v[:, pos]
    return X


def inverse(X):
    v = np.empty((32, 32), complex)                # [lane k2][slot k1]
    for k1 in range(32):
        v[:, k1] = X[lanes + 32 * k1]
    v = pass32(v, np.zeros(32, int))               # slot pos <-> j1 = NBR[pos]
    buf = np.empty_like(v)
    buf[BR[:, None], lanes[None, :]] = v.T         # row brev5(pos), col k2   (same store as the forward transpose)
    w = buf[(-lanes) % 32, :]                      # lane j1 reads row -j1 mod 32: w[j1][k2]
    v = pass32(w, -4 * lanes)                      # slot pos <-> j2 = NBR[pos]
    z = np.empty(1024, complex)
    for pos in range(32):
        j = lanes + 32 * NBR[pos]
        z[j] = v[:, pos] * zeta(-j) / 1024
    return z


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    z = rng.standard_normal(1024) + 1j * rng.standard_normal(1024)
    X = forward(z)
    k = np.arange(1024)
    ref = np.array([np.sum(z * zeta(np.arange(1024) * (4 * kk + 1))) for kk in k])
    print("forward err", np.abs(X - ref).max())
    print("roundtrip err", np.abs(inverse(X) - z).max())
    # negacyclic product check
    a = rng.integers(-100, 100, 2048); b = rng.integers(-100, 100, 2048)
    fa = forward(a[:1024] + 1j * a[1024:]); fb = forward(b[:1024] + 1j * b[1024:])
    c = inverse(fa * fb)
    full = np.convolve(a, b)
    neg = full[:2048].copy(); neg[:2047] -= full[2048:]
    print("negacyclic err", np.abs(np.concatenate([c.real, c.imag]) - neg).max())
    # tangent-form safety: cos != 0 for levels >= 2 in all four tables
    worst = 1.0
    for gs in (np.full(32, 32), 4 * lanes + 1, np.zeros(32, int), -4 * lanes):
        for ci in range(1, 16):
            for g in gs:
                worst = min(worst, abs(np.cos(2 * np.pi * node_exponent(ci, int(g)) / ZN)))
    print("min |cos| over tangent-form constants", worst)
