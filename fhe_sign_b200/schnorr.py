"""Host-side mirror of the reference's Schnorr::sign_fhe_with_k0 (src/schnorr.rs:235-290).

Only the scalar expression `k + e * d` (src/schnorr.rs:272-276) runs under encryption, on the GPU
through fhe_sign_b200.biguint.BigUintFHE; everything else is the reference's plaintext BIP-340 flow
(public key with even y, R = k0 G, parity fix of k, tagged-hash challenge, final `% n` after
decryption), restated here in plain Python integers so that the mirror is self-contained.
"""
import hashlib

from .biguint import BigUintFHE

P = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141        # src/scalar.rs:8
G = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
     0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)


def _add(a, b):
    if a is None or b is None:
        return b if a is None else a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, P) % P
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, P) % P
    x = (lam * lam - a[0] - b[0]) % P
    return (x, (lam * (a[0] - x) - a[1]) % P)


def _mul(k, pt=G):
    acc = None
    while k:
        if k & 1:
            acc = _add(acc, pt)
        pt = _add(pt, pt)
        k >>= 1
    return acc


def _tagged(tag, msg):
    t = hashlib.sha256(tag).digest()
    return hashlib.sha256(t + t + msg).digest()


def _b32(v):
    return int(v).to_bytes(32, "big")


def get_public_key_with_even_y(privkey):
    p = _mul(privkey)
    return p if p[1] % 2 == 0 else (p[0], P - p[1])


def compute_challenge(r, pubkey, message):
    return int.from_bytes(_tagged(b"BIP0340/challenge", _b32(r[0]) + _b32(pubkey[0]) + message), "big") % N


def compute_nonce(privkey, pubkey, message, aux_rand):
    """src/schnorr.rs:394-401 (BIP-340 nonce from the un-negated key, as the reference computes it)."""
    t = bytes(x ^ y for x, y in zip(_b32(privkey), _tagged(b"BIP0340/aux", aux_rand)))
    return int.from_bytes(_tagged(b"BIP0340/nonce", t + _b32(pubkey[0]) + message), "big") % N


class Signature:
    def __init__(self, r_x, s):
        self.r_x, self.s = r_x, s

    def to_bytes(self):
        return _b32(self.r_x) + _b32(self.s)


def sign_fhe_with_k0(message, k0, privkey, privkey_fhe, client_key, fused=False, reduce_encrypted=False, public_challenge=False):
    """src/schnorr.rs:235-290.  `privkey` is the plaintext key (used only for the public key, :241);
    `privkey_fhe` is its BigUintFHE encryption.  fused=True evaluates the same expression with the
    batched schedule (BigUintFHE.mul_add_fused).  reduce_encrypted=True also takes `mod n` under encryption
    (SURVEY.md 8f.2), so that only the 256-bit s is ever decrypted; the reference reduces in plaintext (:276).
    public_challenge=True is a protocol-level variant, NOT the reference's dataflow: the challenge e = H(R || P || m) is public by
    construction (verify recomputes it, :307-345), so it is handed to the server in plaintext and k + e*d becomes a
    scalar-times-ciphertext product (fsc_radix_scalar_mul_add_wide); d and the nonce k stay encrypted.  Same signature bytes."""
    pubkey = get_public_key_with_even_y(privkey)
    r = _mul(k0)
    k = N - k0 if r[1] % 2 == 1 else k0
    e = compute_challenge(r, pubkey, message)
    k_fhe = BigUintFHE.new(k, client_key)
    if public_challenge:
        s_fhe = BigUintFHE.scalar_mul_add_fused(k_fhe, e, privkey_fhe.clone())
        if reduce_encrypted:
            s_fhe = s_fhe.rem_scalar(N)
        return Signature(r[0], s_fhe.to_biguint(client_key) % N)
    e_fhe = BigUintFHE.new(e, client_key)
    if fused:
        s_fhe = BigUintFHE.mul_add_fused(k_fhe, e_fhe, privkey_fhe.clone())
    else:
        s_fhe = k_fhe + (e_fhe * privkey_fhe.clone())
    if reduce_encrypted:
        s_fhe = s_fhe.rem_scalar(N)
    s_without_mod = s_fhe.to_biguint(client_key)
    return Signature(r[0], s_without_mod % N)


def sign_fhe(message, aux_rand, privkey, client_key, fused=False, public_challenge=False):
    """src/schnorr.rs:154-208: the nonce is derived from aux_rand in plaintext (:168), the private key is encrypted here
    (:192: `BigUintFHE::new(privkey)`), then the same scalar expression as sign_fhe_with_k0 (:195 = :274).
    Known answer (src/schnorr.rs:440-466): vector 0 -> E907831F...310536C0."""
    pubkey = get_public_key_with_even_y(privkey)
    k0 = compute_nonce(privkey, pubkey, message, aux_rand)
    return sign_fhe_with_k0(message, k0, privkey, BigUintFHE.new(privkey, client_key), client_key, fused=fused,
                            public_challenge=public_challenge)
