"""Minimal driver for ncu: a few launches of the keyswitch and PBS kernels on synthetic data."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fhe_sign_b200 as fsb
from fhe_sign_b200.capi import LWE_BIG, LWE_SMALL

count = int(sys.argv[1]) if len(sys.argv) > 1 else 592
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
acc_bits = int(os.environ.get("FSC_BENCH_ACC_BITS", "32"))
p = fsb.Params.preset("2_2_gaussian", acc_bits=acc_bits)
n = p.lwe_dim
rng = np.random.default_rng(1)
ctx = fsb.Context(p)
ctx.upload_keys(rng.integers(0, 2**64, n * 4 * 2048, dtype=np.uint64), rng.integers(0, 2**64, 2048 * 5 * (n + 1), dtype=np.uint64))
luts = ctx.luts_from_tables(np.arange(16))
din = ctx.lwe(LWE_BIG, count).upload(rng.integers(0, 2**64, (count, 2049), dtype=np.uint64))
dsm, dout = ctx.lwe(LWE_SMALL, count), ctx.lwe(LWE_BIG, count)
for _ in range(reps):
    ctx.keyswitch(din, dsm)
    ctx.timer_start()
    ctx.pbs(dsm, luts, None, dout)
    ms = ctx.timer_stop()
    print("pbs %d cts: %.3f ms -> %.0f PBS/s" % (count, ms, count / ms * 1e3))
ctx.close()
