"""Operator latency bench (BASELINE.json configs[2..4] + the perf_test.rs op list): one radix operator at
a time on synthetic random ciphertext blocks (cost is data independent; correctness is the tests' job).
Prints one JSON line per operator: wall-clock ms including host scheduling, PBS count, PBS levels."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fhe_sign_b200 as fsb

N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


def main():
    preset = os.environ.get("FSC_BENCH_PRESET", "2_2_gaussian")
    acc_bits = int(os.environ.get("FSC_BENCH_ACC_BITS", "32"))
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    p = fsb.Params.preset(preset, acc_bits=acc_bits)
    n = p.lwe_dim
    rng = np.random.default_rng(3)
    ctx = fsb.Context(p)
    ctx.upload_keys(rng.integers(0, 2**64, n * 4 * 2048, dtype=np.uint64), rng.integers(0, 2**64, 2048 * 5 * (n + 1), dtype=np.uint64))
    R = ctx.radix

    def rnd(blocks):
        return R.from_lwe(rng.integers(0, 2**64, (blocks, 2049), dtype=np.uint64))

    a8, b8 = rnd(4), rnd(4)
    a32, b32 = rnd(16), rnd(16)
    a64z, b64z = R.cast(rnd(16), 32), R.cast(rnd(16), 32)       # zero-extended u32 -> u64, as biguint.rs does
    a256, b256, c256 = rnd(128), rnd(128), rnd(128)
    a513 = rnd(257)
    ops = [
        ("u32 add (perf_test.rs:28)", lambda: a32 + b32),
        ("u32 mul (perf_test.rs:32)", lambda: a32 * b32),
        ("u32 shr by encrypted amount (perf_test.rs:36)", lambda: a32 >> b32),
        ("u8 min (perf_test.rs:44)", lambda: R.min(a8, b8)),
        ("u8 and 1 (perf_test.rs:48)", lambda: a8 & 1),
        ("u32 div 5 (perf_test.rs:54)", lambda: a32 // 5),
        ("u64 add of zero-extended u32 (biguint.rs:138)", lambda: a64z + b64z),
        ("u64 mul of zero-extended u32 (biguint.rs:223)", lambda: a64z * b64z),
        ("256-bit add", lambda: a256 + b256),
        ("256-bit mul (wrapping, 128 blocks)", lambda: a256 * b256),
        ("256-bit shr by scalar 77", lambda: a256 >> 77),
        ("256-bit shr by encrypted amount", lambda: a256 >> b256),
        ("256x256->512-bit mul (mul_wide)", lambda: R.mul_wide(a256, b256, 256)),
        ("k + e*d fused (schnorr.rs:274)", lambda: R.mul_add_wide(a256, b256, c256, 272)),
        ("256-bit div 5", lambda: a256 // 5),
        ("514-bit rem n (secp256k1 order)", lambda: a513 % N_ORDER),
    ]
    for name, fn in ops:
        best = None
        for _ in range(reps):
            ctx.sync()
            p0, l0 = R.stats()
            t0 = time.perf_counter()
            out = fn()
            ctx.sync()
            dt = (time.perf_counter() - t0) * 1e3
            p1, l1 = R.stats()
            del out
            best = dt if best is None else min(best, dt)
        print(json.dumps({"op": name, "ms": round(best, 3), "pbs": p1 - p0, "levels": l1 - l0,
                          "ms_per_level": round(best / max(l1 - l0, 1), 3), "acc_bits": acc_bits, "preset": preset}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
