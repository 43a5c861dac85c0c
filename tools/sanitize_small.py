"""Small driver for compute-sanitizer (memcheck / racecheck): every kernel of the hot path once, on the toy parameter set
(n = 48, same GLWE side) and batch widths that select each blind-rotation kernel, results checked by decryption.

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_sign_b200 as fsb  # noqa: E402
from fhe_sign_b200.capi import LWE_BIG  # noqa: E402
from oracle import orc  # noqa: E402


def main():
    K = orc.Keys(orc.preset("toy"), 1)
    table = (np.arange(16) * 5 + 2) % 16
    rng = np.random.default_rng(2)
    for variant, acc_bits, counts in (("auto", 32, (5, 160, 310)), ("auto", 64, (5, 310)), ("solo", 32, (310,)), ("quad", 32, (310, 593)), ("duo", 32, (310, 593))):
        os.environ["FSC_PBS_VARIANT"] = variant
        ctx = fsb.Context(fsb.Params.preset("toy", acc_bits=acc_bits))
        ctx.upload_keys(K.bsk, K.ksk)                      # bsk_exact_kernel, ksk_limb_transpose_kernel
        luts = ctx.luts_from_tables(table)
        for count in counts:
            m = rng.integers(0, 16, count).astype(np.uint64)
            din, dout = ctx.lwe(LWE_BIG, count).upload(K.encrypt_msgs(m)), ctx.lwe(LWE_BIG, count)
            ctx.ks_pbs(din, luts, None, dout)              # ks_decompose_kernel, ks_umma_kernel, pbs_{split,stream,ring,solo,quad,duo}_kernel
            ok = (K.decrypt_msgs(dout.download()) == table[m]).all()
            print("%s acc %d count %d kernel %s: %s" % (variant, acc_bits, count, ctx.pbs_kernel_name(), "ok" if ok else "WRONG"), flush=True)
            assert ok
            din.free(); dout.free()
        if variant == "auto" and acc_bits == 32:           # radix layer: lincomb_kernel, scatter / gather
            R = ctx.radix
            digits = lambda v, n: np.array([(v >> (2 * i)) & 3 for i in range(n)], dtype=np.uint64)
            a, b = R.from_lwe(K.encrypt_msgs(digits(1344, 16), seed=3)), R.from_lwe(K.encrypt_msgs(digits(5, 16), seed=4))
            d = K.decrypt_msgs(R.to_lwe(a * b))
            assert sum(int(v) << (2 * i) for i, v in enumerate(d)) == 6720
            print("radix u32 mul: ok", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
