"""256-bit radix operators with PBS levels sharded over the GPUs of one node (BASELINE configs[2..4]).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_ops.py
             [--exchange peer|nccl] [--quick]
Every rank holds replicated keys and runs the same operator sequence; wide levels are sliced across ranks.
  --exchange peer (default): the library's own exchange - block pools mapped into every peer, the blind rotation's
                  epilogue stores its outputs into all pools over NVLink, one flag-barrier kernel per level (fsc_peer_pool_*);
  --exchange nccl: the callback form - NCCL all-gather enqueued from Python per level + a scatter kernel (baseline).
Rank 0 checks decrypted results with the oracle client (seeded keys) and prints one JSON line per operator (time = max
over ranks)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fhe_sign_b200 as fsb
from fhe_sign_b200.distributed import enable_level_sharding, enable_peer_sharding
from oracle import orc
from oracle_client import OracleClientKey

N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    exchange = sys.argv[sys.argv.index("--exchange") + 1] if "--exchange" in sys.argv else "peer"
    quick = "--quick" in sys.argv
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K = orc.Keys(orc.preset("2_2_gaussian"), 1)
    stream = torch.cuda.Stream()
    ctx = fsb.Context(fsb.Params.preset("2_2_gaussian", acc_bits=32), device=local, stream=stream.cuda_stream)
    ctx.upload_keys(K.bsk, K.ksk)
    min_width = int(os.environ.get("FSC_SHARD_MIN", "149"))
    if world > 1:
        if exchange == "peer":
            enable_peer_sharding(ctx, min_width=min_width)
        else:
            enable_level_sharding(ctx, stream, min_width=min_width, capacity_blocks=1 << 16)
    R = ctx.radix
    ck = OracleClientKey(K, seed=77)          # same seed on every rank: identical ciphertexts everywhere
    rnd = np.random.default_rng(5)
    x, y, z = (int.from_bytes(rnd.bytes(32), "little") for _ in range(3))
    a, b, c = ck.encrypt_blocks(x, 128, R), ck.encrypt_blocks(y, 128, R), ck.encrypt_blocks(z, 128, R)
    w = ck.encrypt_blocks(x * y + z, 257, R)
    ops = [
        ("256-bit mul (wrapping)", lambda: a * b, (x * y) % 2**256),
        ("256-bit shr by encrypted amount", lambda: a >> b, x >> (y % 256)),
        ("k + e*d fused, 256x256+256", lambda: R.mul_add_wide(a, b, c, 272), x * y + z),
        ("256-bit div 5", lambda: a // 5, x // 5),
        ("514-bit rem n", lambda: w % N_ORDER, (x * y + z) % N_ORDER),
    ]
    if quick:
        ops = ops[:2] + [("256-bit min", lambda: R.min(a, b), min(x, y))]
    for name, fn, want in ops:
        for rep in range(1 if quick else 2):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            p0, l0 = R.stats(); s0 = R.sharded_levels()
            t0 = time.perf_counter()
            out = fn()
            ctx.sync(); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        p1, l1 = R.stats()
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        got = ck.decrypt(out, R)
        ok = got == want
        if rank == 0:
            print(json.dumps({"op": name, "n_gpus": world, "exchange": exchange if world > 1 else "none", "ms": round(float(t[0]) * 1e3, 2),
                              "pbs": p1 - p0, "levels": l1 - l0, "sharded_levels": R.sharded_levels() - s0, "correct": bool(ok)}), flush=True)
        assert ok, (name, rank)
    if world > 1:
        ctx.sync()
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
