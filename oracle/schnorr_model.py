"""Plaintext restatement of the reference's signing dataflow — TEST INFRASTRUCTURE (oracle).

Follows, line by line:
  src/secp256k1.rs            affine point add / double / scalar_mul, generator
  src/schnorr.rs:114-141      Schnorr::sign_with_k0
  src/schnorr.rs:235-290      Schnorr::sign_fhe_with_k0 (the FHE section :272-276 replaced by the
                              digit-level replay of src/biguint.rs below)
  src/schnorr.rs:352-410      get_public_key_with_even_y, tagged_hash, compute_nonce, compute_challenge
  src/biguint.rs:120-192      impl Add for BigUintFHE  (ripple carry over u32 digits through u64 sums)
  src/biguint.rs:194-265      impl Mul for BigUintFHE  (schoolbook; the carry out of result[idx+2] is
                              dropped by a wrapping FheUint32 add, :247-249)
The reference never negates d for an odd-y public key, and neither does this.
"""
import hashlib

P = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
M32 = 0xFFFFFFFF


def point_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0] and (a[1] + b[1]) % P == 0:
        return None
    if a == b:
        lam = 3 * a[0] * a[0] * pow(2 * a[1], P - 2, P) % P
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], P - 2, P) % P
    x = (lam * lam - a[0] - b[0]) % P
    return (x, (lam * (a[0] - x) - a[1]) % P)


def scalar_mul(k, pt=(GX, GY)):
    r = None
    while k:
        if k & 1:
            r = point_add(r, pt)
        pt = point_add(pt, pt)
        k >>= 1
    return r


def tagged_hash(tag, msg):
    t = hashlib.sha256(tag).digest()
    return hashlib.sha256(t + t + msg).digest()


def b32(v):
    return int(v).to_bytes(32, "big")


def pubkey_even_y(d):
    p = scalar_mul(d)
    return p if p[1] % 2 == 0 else (p[0], P - p[1])


def compute_nonce(d, pub, message, aux):
    t = bytes(x ^ y for x, y in zip(b32(d), tagged_hash(b"BIP0340/aux", aux)))
    return int.from_bytes(tagged_hash(b"BIP0340/nonce", t + b32(pub[0]) + message), "big") % N


def compute_challenge(r, pub, message):
    return int.from_bytes(tagged_hash(b"BIP0340/challenge", b32(r[0]) + b32(pub[0]) + message), "big") % N


def signing_inputs(message, k0, d):
    """plaintext part of sign_with_k0 / sign_fhe_with_k0 up to the scalar expression: (R, k, e)."""
    pub = pubkey_even_y(d)
    r = scalar_mul(k0)
    k = N - k0 if r[1] % 2 == 1 else k0
    return r, k, compute_challenge(r, pub, message)


def sign_with_k0(message, k0, d):
    r, k, e = signing_inputs(message, k0, d)
    return b32(r[0]) + b32((k + e * d) % N)


def verify(message, pub_x, sig):
    if pub_x >= P:
        return False
    y2 = (pow(pub_x, 3, P) + 7) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        return False
    if y % 2:
        y = P - y
    r, s = int.from_bytes(sig[:32], "big"), int.from_bytes(sig[32:], "big")
    if r >= P or s >= N:
        return False
    e = int.from_bytes(tagged_hash(b"BIP0340/challenge", sig[:32] + b32(pub_x) + message), "big") % N
    pt = point_add(scalar_mul(s), scalar_mul(N - e, (pub_x, y)))
    return pt is not None and pt[1] % 2 == 0 and pt[0] == r


# ---- digit-level replay of BigUintFHE --------------------------------------------------------
def to_u32_digits(v):
    d = []
    while v:
        d.append(v & M32)
        v >>= 32
    return d


def from_digits(d):
    return sum(x << (32 * i) for i, x in enumerate(d))


def biguint_add(a, b):
    """src/biguint.rs:123-191 on plaintext u32 digit lists."""
    res, carry = [], None
    for i in range(max(len(a), len(b))):
        x = a[i] if i < len(a) else None
        y = b[i] if i < len(b) else None
        if x is not None and y is not None and carry is not None:
            t = x + y + carry
        elif x is not None and y is not None:
            t = x + y
        elif x is not None and carry is not None:
            t = x + carry
        elif x is not None:
            res.append(x); continue
        elif y is not None and carry is not None:
            t = y + carry
        elif y is not None:
            res.append(y); continue
        else:
            res.append(carry); continue
        carry = (t >> 32) & M32
        res.append(t & M32)
    if carry is not None:
        res.append(carry)
    return res


def biguint_mul(a, b):
    """src/biguint.rs:197-264 on plaintext u32 digit lists, including the dropped carry at :247-249."""
    if not a or not b:
        return []
    res = [0] * (len(a) + len(b))
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            idx = i + j
            prod = x * y
            lower, upper = prod & M32, prod >> 32
            s = res[idx] + lower
            res[idx] = s & M32
            s = res[idx + 1] + upper + (s >> 32)
            res[idx + 1] = s & M32
            if idx + 2 < len(res):
                res[idx + 2] = (res[idx + 2] + (s >> 32)) & M32        # wrapping FheUint32 add
    return res


def sign_fhe_with_k0_model(message, k0, d):
    """what the reference's FHE path decrypts to: (signature bytes, k + e*d before the plaintext mod n)."""
    r, k, e = signing_inputs(message, k0, d)
    s_digits = biguint_add(to_u32_digits(k), biguint_mul(to_u32_digits(e), to_u32_digits(d)))
    s_wo = from_digits(s_digits)
    return b32(r[0]) + b32(s_wo % N), s_wo
