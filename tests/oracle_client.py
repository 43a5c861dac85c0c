"""ClientKey stand-in backed by the oracle (test infrastructure): FheUint32::try_encrypt / decrypt
(src/biguint.rs:26,70) performed on the CPU with the seeded oracle keys."""
import numpy as np


class OracleClientKey:
    def __init__(self, K, seed=4242):
        self.K, self.seed, self.stream = K, seed, 0

    def encrypt_blocks(self, value, n_blocks, api):
        digits = np.array([(int(value) >> (2 * i)) & 3 for i in range(n_blocks)], dtype=np.uint64)
        ct = self.K.encrypt_msgs(digits, seed=self.seed, stream=self.stream)
        self.stream += n_blocks
        return api.from_lwe(ct)

    def encrypt_u32(self, value, api):
        return self.encrypt_blocks(value, 16, api)

    def decrypt(self, r, api):
        d = self.K.decrypt_msgs(api.to_lwe(r))
        assert (d < 4).all(), "decrypted block carries are not empty: %s" % d
        return sum(int(v) << (2 * i) for i, v in enumerate(d))
