"""One 256-bit add and one u32 add on synthetic blocks: the kernels of a narrow PBS level, for an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/level_profile.py)."""
import os
import sys
import time

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
a256, b256, a32, b32 = rnd(128), rnd(128), rnd(16), rnd(16)
for name, fn in (("256-bit add", lambda: a256 + b256), ("u32 add", lambda: a32 + b32)):
    for rep in range(2):
        ctx.sync()
        p0, l0 = R.stats()
        t0 = time.perf_counter()
        out = fn()
        t_host = (time.perf_counter() - t0) * 1e3
        ctx.sync()
        dt = (time.perf_counter() - t0) * 1e3
        p1, l1 = R.stats()
    print("%s: %.2f ms (%d levels, %.3f ms per level; host enqueue done after %.2f ms), %d PBS" % (name, dt, l1 - l0, dt / (l1 - l0), t_host, p1 - p0))
ctx.close()
