"""BigUintFHE — host-side mirror of the reference's src/biguint.rs over the GPU radix engine.

Same public surface and the same digit-level dataflow as the reference (little-endian Vec<FheUint32>,
variable length, ripple-carry Add through FheUint64 sums, schoolbook Mul including the wrapping add
into result[idx + 2], src/biguint.rs:120-265), with every FheUint32/FheUint64 operator forwarded to
libfhe_sign_cuda.so instead of the tfhe-rs CPU server key.  `set_server_key` plays the role of
tfhe::set_server_key (src/biguint.rs:278): it installs the GPU context the operators use.

Two schedules:
  * faithful (`+`, `*`): op for op what biguint.rs does — the parity mode;
  * fused (`mul_add_fused`): the same mathematical result for k + e*d computed as ONE 128-block radix
    product + one multi-operand sum, which is what exposes wide PBS batches to the GPU.
    It equals the faithful result whenever the reference itself does not drop a carry (all BIP-340 vectors).
"""
import threading

_tls = threading.local()

BLOCKS_U32, BLOCKS_U64 = 16, 32
M32 = 0xFFFFFFFF


def set_server_key(ctx):
    """Installs the GPU server key (an fhe_sign_b200.Context with keys uploaded) for this thread."""
    _tls.api = ctx.radix


def _api():
    api = getattr(_tls, "api", None)
    if api is None:
        raise RuntimeError("no server key installed: call fhe_sign_b200.biguint.set_server_key(ctx) first")
    return api


def cast64(x):
    """FheUint64::cast_from(FheUint32) — src/biguint.rs:135-137,221-222 (zero extension, no PBS)."""
    return _api().cast(x, BLOCKS_U64)


def cast32(x):
    """FheUint32::cast_from(FheUint64) — truncation, no PBS."""
    return _api().cast(x, BLOCKS_U32)


class BigUintFHE:
    def __init__(self, digits, client_key):
        self.digits = list(digits)            # little-endian FheUint32 look-alikes (16-block radix values)
        self.client_key = client_key

    # ---- constructors (src/biguint.rs:17-58) -----------------------------------------------
    @classmethod
    def new(cls, value, client_key):
        value = int(value)
        if value == 0:
            return cls([], client_key)
        digits = []
        while value:
            digits.append(value & M32)
            value >>= 32
        return cls([client_key.encrypt_u32(d, _api()) for d in digits], client_key)

    @classmethod
    def from_u32(cls, value, client_key):
        return cls.new(value & M32, client_key)

    @classmethod
    def from_encrypted_digits(cls, digits, client_key):
        return cls(digits, client_key)

    @classmethod
    def zero(cls, client_key):
        return cls([], client_key)

    @classmethod
    def one(cls, client_key):
        return cls.from_u32(1, client_key)

    def clone(self):
        return BigUintFHE(self.digits, self.client_key)      # radix values are immutable: sharing is a deep copy

    # ---- decryption (src/biguint.rs:61-105) --------------------------------------------------
    def to_biguint(self, client_key):
        return sum(client_key.decrypt(d, _api()) << (32 * i) for i, d in enumerate(self.digits))

    def decrypt_to_u32(self, client_key):
        if len(self.digits) == 0:
            return 0
        if len(self.digits) == 1:
            return client_key.decrypt(self.digits[0], _api())
        return None

    def decrypt_to_u64(self, client_key):
        if len(self.digits) > 2:
            return None
        return self.to_biguint(client_key)

    # ---- src/biguint.rs:108-117 ----------------------------------------------------------------
    @staticmethod
    def extract_upper_bits(s64):
        return cast32(s64 >> 32)

    @staticmethod
    def extract_lower_bits(s64):
        return cast32(s64 & M32)

    # ---- impl Add (src/biguint.rs:120-192) -------------------------------------------------------
    def __add__(self, other):
        result, carry = [], None
        for i in range(max(len(self.digits), len(other.digits))):
            a = self.digits[i] if i < len(self.digits) else None
            b = other.digits[i] if i < len(other.digits) else None
            if a is not None and b is not None and carry is not None:
                t = cast64(a) + cast64(b) + cast64(carry)
            elif a is not None and b is not None:
                t = cast64(a) + cast64(b)
            elif a is not None and carry is not None:
                t = cast64(a) + cast64(carry)
            elif a is not None:
                result.append(a); continue
            elif b is not None and carry is not None:
                t = cast64(b) + cast64(carry)
            elif b is not None:
                result.append(b); continue
            else:
                result.append(carry); continue
            carry = cast32(t >> 32)
            result.append(cast32(t & M32))
        if carry is not None:
            result.append(carry)
        return BigUintFHE(result, self.client_key)

    # ---- impl Mul (src/biguint.rs:194-265) ---------------------------------------------------------
    def __mul__(self, other):
        if not self.digits or not other.digits:
            return BigUintFHE([], self.client_key)
        zero = self.client_key.encrypt_u32(0, _api())                        # :206-209, one encryption cloned
        result = [zero] * (len(self.digits) + len(other.digits))
        for i, a in enumerate(self.digits):
            for j, b in enumerate(other.digits):
                idx = i + j
                product = cast64(a) * cast64(b)                              # :221-223
                lower = BigUintFHE.extract_lower_bits(product)
                upper = BigUintFHE.extract_upper_bits(product)
                s = cast64(result[idx]) + cast64(lower)                      # :234-236
                result[idx] = BigUintFHE.extract_lower_bits(s)
                s2 = cast64(result[idx + 1]) + cast64(upper) + cast64(BigUintFHE.extract_upper_bits(s))   # :240-243
                result[idx + 1] = BigUintFHE.extract_lower_bits(s2)
                if idx + 2 < len(result):                                    # :247-249 (wrapping u32 add)
                    result[idx + 2] = result[idx + 2] + BigUintFHE.extract_upper_bits(s2)
        return BigUintFHE(result, self.client_key)

    # ---- fused schedule --------------------------------------------------------------------------------
    def _as_radix(self):
        api = _api()
        return api.concat(self.digits) if self.digits else api.trivial(0, 0)

    @staticmethod
    def _from_radix(r, n_digits, client_key):
        api = _api()
        return BigUintFHE([api.slice(r, BLOCKS_U32 * i, BLOCKS_U32) for i in range(n_digits)], client_key)

    @staticmethod
    def mul_add_fused(k, e, d):
        """k + e*d with the digit-count conventions of `k + (e * d)` in the reference (src/schnorr.rs:274)."""
        api = _api()
        if not e.digits or not d.digits:
            return k + BigUintFHE([], k.client_key)
        n_prod = len(e.digits) + len(d.digits)
        n_out = max(len(k.digits), n_prod) + 1                               # Add appends the final carry digit
        # the addend joins the product's column sum: one carry propagation for both (e*d < 2^(32 n_prod), so nothing is lost
        # by forming the sum at n_out digits directly)
        total = api.mul_add_wide(e._as_radix(), d._as_radix(), k._as_radix(), BLOCKS_U32 * n_out)
        return BigUintFHE._from_radix(total, n_out, k.client_key)

    @staticmethod
    def scalar_mul_add_fused(k, e_plain, d):
        """k + e*d with the challenge e in PLAINTEXT (public by construction: every verifier recomputes it), d and k encrypted:
        fsc_radix_scalar_mul_add_wide.  Same digit-count conventions as mul_add_fused."""
        api = _api()
        e_plain = int(e_plain)
        if e_plain == 0 or not d.digits:
            return k + BigUintFHE([], k.client_key)
        n_prod = (e_plain.bit_length() + 31) // 32 + len(d.digits)
        n_out = max(len(k.digits), n_prod) + 1
        total = api.scalar_mul_add_wide(d._as_radix(), e_plain, k._as_radix(), BLOCKS_U32 * n_out)
        return BigUintFHE._from_radix(total, n_out, k.client_key)

    def rem_scalar(self, modulus):
        """self mod a plaintext modulus, homomorphically (the reference takes `% n` after decryption,
        src/schnorr.rs:276).  Digits follow the modulus' width."""
        api = _api()
        if not self.digits:
            return BigUintFHE([], self.client_key)
        n_digits = max(1, (int(modulus).bit_length() + 31) // 32)
        r = api.scalar("rem", self._as_radix(), int(modulus))
        return BigUintFHE._from_radix(r, min(n_digits, len(self.digits)), self.client_key)
