"""Per-level device times of the fused k + e*d (FSC_LEVEL_TRACE=1) on synthetic blocks."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fhe_sign_b200 as fsb
p = fsb.Params.preset("2_2_gaussian", acc_bits=32)
n = p.lwe_dim
rng = np.random.default_rng(3)
ctx = fsb.Context(p)
ctx.upload_keys(rng.integers(0, 2**64, n * 4 * 2048, dtype=np.uint64), rng.integers(0, 2**64, 2048 * 5 * (n + 1), dtype=np.uint64))
R = ctx.radix
rnd = lambda blocks: R.from_lwe(rng.integers(0, 2**64, (blocks, 2049), dtype=np.uint64))
a, b, c = rnd(128), rnd(128), rnd(128)
for rep in range(2):
    ctx.sync(); t0 = time.perf_counter()
    print("--- rep", rep, file=sys.stderr, flush=True)
    out = R.mul_add_wide(a, b, c, 272)
    ctx.sync()
    print("k + e*d: %.1f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
ctx.close()
