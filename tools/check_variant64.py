"""Quick decrypt check at the 64-bit accumulator (default variant) on the toy parameter set."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fhe_sign_b200 as fsb
from fhe_sign_b200.capi import LWE_BIG
from oracle import orc
K = orc.Keys(orc.preset("toy"), 1)
table = (np.arange(16) * 5 + 2) % 16
rng = np.random.default_rng(2)
ctx = fsb.Context(fsb.Params.preset("toy", acc_bits=64))
ctx.upload_keys(K.bsk, K.ksk)
luts = ctx.luts_from_tables(table)
for count in [int(a) for a in sys.argv[1:]] or (5, 200, 310, 1200):
    m = rng.integers(0, 16, count).astype(np.uint64)
    din, dout = ctx.lwe(LWE_BIG, count).upload(K.encrypt_msgs(m)), ctx.lwe(LWE_BIG, count)
    ctx.ks_pbs(din, luts, None, dout)
    got = K.decrypt_msgs(dout.download())
    print("acc64 count %d kernel %s: %d wrong" % (count, ctx.pbs_kernel_name(), int((got != table[m]).sum())), flush=True)
    assert (got == table[m]).all()
ctx.close()
