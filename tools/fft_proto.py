"""numpy prototype of the warp-level 32x32 negacyclic FFT used by the PBS kernel
(fhe_sign_b200/csrc/pbs_core.cuh).  Index conventions here are the reference for the CUDA code.

N = 2048 real coefficients -> M = 1024 complex z_j = a_j + i a_{j+M};
X_k = sum_j z_j zeta^{j(4k+1)}, zeta = exp(2 pi i / 4096).
j = j1 + 32 j2 (pass 1: lane j1, register j2), k = k2 + 32 k1 (pass 2: lane k2).
Each pass evaluates a degree-<32 polynomial at the 32 roots of x^32 = zeta^(32 g)
(g = 32 in pass 1, g = 4*lane+1 in pass 2) by recursive splitting x^2h - c = (x^h - s)(x^h + s).
"""
import numpy as np

ZN = 4096


def zeta(e):
    e = np.asarray(e) % ZN
    return np.exp(2j * np.pi * e / ZN)


def u_exponents():
    """exponent (units of 2pi/64) of the uniform factor u_{L,m}, L=1..5, m<2^(L-1)."""
    u = {1: [0]}
    for L in range(1, 5):
        nxt = []
        for e in u[L]:
            nxt += [e // 2, e // 2 + 16]
        u[L + 1] = nxt
    return u


U = u_exponents()
# register position -> index (k2 in pass 1, k1 in pass 2) of the root held there
POS2IDX = []
for m, e in enumerate(U[5]):
    for sign in (0, 1):
        POS2IDX.append((e // 2 + 16 * sign) % 32)
POS2IDX = np.array(POS2IDX)
IDX2POS = np.argsort(POS2IDX)


def node_consts(g):
    """s_{L,m} = zeta^{(32>>L) g} * omega_64^{u_{L,m}} for a vector of g (one per lane)."""
    g = np.asarray(g)
    out = {}
    for L in range(1, 6):
        out[L] = [zeta((32 >> L) * g + 64 * e) for e in U[L]]      # omega_64 = zeta^64
    return out


def dft32_fwd(v, S):
    """v: [lanes, 32] complex, S: node constants (per lane arrays). In place CT butterflies."""
    for L in range(1, 6):
        half = 16 >> (L - 1)
        for m in range(1 << (L - 1)):
            base = m * 2 * half
            s = S[L][m][:, None] if np.ndim(S[L][m]) else S[L][m]
            lo = v[:, base:base + half].copy(); hi = v[:, base + half:base + 2 * half] * s
            v[:, base:base + half] = lo + hi
            v[:, base + half:base + 2 * half] = lo - hi
    return v


def dft32_inv(v, S):
    """exact reverse of dft32_fwd up to a factor 32 (GS butterflies with conj constants)."""
    for L in range(5, 0, -1):
        half = 16 >> (L - 1)
        for m in range(1 << (L - 1)):
            base = m * 2 * half
            s = S[L][m][:, None] if np.ndim(S[L][m]) else S[L][m]
            u = v[:, base:base + half].copy(); w = v[:, base + half:base + 2 * half].copy()
            v[:, base:base + half] = u + w
            v[:, base + half:base + 2 * half] = (u - w) * np.conj(s)
    return v


lanes = np.arange(32)
S1 = node_consts(np.full(32, 32))
S2 = node_consts(4 * lanes + 1)


def fwd(a):
    """a: 2048 reals -> X as [lane k2, pos] (frequency k = k2 + 32*POS2IDX[pos])."""
    z = a[:1024] + 1j * a[1024:]
    regs = z.reshape(32, 32).T.copy()          # regs[j1, j2] = z[j1 + 32 j2]
    dft32_fwd(regs, S1)                        # regs[j1, pos] = Y[j1][k2 = POS2IDX[pos]]
    xbuf = np.empty((32, 32), complex)
    xbuf[POS2IDX[None, :], lanes[:, None]] = regs      # xbuf[k2][j1]
    regs2 = xbuf.copy()                        # lane k2 reads row k2: regs2[k2, j1]
    dft32_fwd(regs2, S2)
    return regs2


def inv(X):
    regs2 = X.copy()
    dft32_inv(regs2, S2)                       # regs2[k2, j1] = 32 Y[j1][k2]
    xbuf = regs2                               # xbuf[k2][j1]
    regs = xbuf[POS2IDX[None, :], lanes[:, None]]      # regs[j1, pos] = xbuf[k2(pos)][j1]
    regs = regs.copy()
    dft32_inv(regs, S1)                        # regs[j1, j2] = 1024 z[j1 + 32 j2]
    z = (regs.T.reshape(-1)) / 1024.0
    return np.concatenate([z.real, z.imag])


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    a = rng.standard_normal(2048)
    X = fwd(a)
    # direct definition
    j = np.arange(1024)
    z = a[:1024] + 1j * a[1024:]
    for k2 in (0, 5, 31):
        for pos in (0, 1, 7, 31):
            k = k2 + 32 * POS2IDX[pos]
            ref = np.sum(z * zeta(j * (4 * k + 1)))
            assert abs(ref - X[k2, pos]) < 1e-8, (k2, pos, ref, X[k2, pos])
    assert np.allclose(inv(X), a)
    # negacyclic product
    b = rng.integers(-5, 5, 2048).astype(float)
    c = inv(fwd(a) * fwd(b))
    full = np.convolve(a, b)
    ref = full[:2048].copy(); ref[:2047] -= full[2048:]
    assert np.allclose(c, ref)
    print("ok", POS2IDX.tolist())
    print("U", U)
