"""BASELINE.json configs[4]: full Schnorr sign_fhe_with_k0 over the BIP-340 signing rows (tests/golden/schnorr_vectors.json,
generated from the reference's tests/test_vectors.csv) on N GPUs of one node.

    python tools/sign_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sign_bench.py

One process per GPU, keys replicated (same seeded keygen on every rank), signatures are independent units: row r goes
to rank r mod N, no data-path collective.  Everything runs through the product's own C ABI:
fsc_client_* for keys / encryption / decryption, fsc_radix_* for the evaluation; the oracle is not involved.  Every
signature is compared with the reference's plaintext twin (sign_with_k0) bytes recorded in the golden file.
Prints one JSON line per signature and one summary line (time = max over ranks, as the reference-side clock would see).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_sign_b200 as fsb
from fhe_sign_b200 import biguint as bg
from fhe_sign_b200 import schnorr
from fhe_sign_b200.biguint import BigUintFHE
from fhe_sign_b200.client import generate_keys


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    fused = "--faithful" not in sys.argv
    reps = 2 if "--twice" in sys.argv else 1
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "schnorr_vectors.json")))
    preset = os.environ.get("FSC_BENCH_PRESET", "2_2_gaussian")
    t0 = time.perf_counter()
    ck, (bsk, ksk) = generate_keys(preset, seed=2024)
    t_keygen = time.perf_counter() - t0
    ctx = fsb.Context(fsb.Params.preset(preset, acc_bits=32), device=local)
    t0 = time.perf_counter()
    ctx.upload_keys(bsk, ksk)
    t_upload = time.perf_counter() - t0
    bg.set_server_key(ctx)
    api = bg._api()
    mine = [v for i, v in enumerate(golden) if i % world == rank]
    lines = []
    if dist is not None:
        dist.barrier()
    t_all = time.perf_counter()
    for rep in range(reps):
        for v in mine:
            d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
            p0, l0 = api.stats()
            t0 = time.perf_counter()
            d_fhe = BigUintFHE.new(d, ck)
            sig = schnorr.sign_fhe_with_k0(msg, k0, d, d_fhe, ck, fused=fused)
            dt = time.perf_counter() - t0
            p1, l1 = api.stats()
            ok = sig.to_bytes().hex().upper() == v["reference_signature"]
            lines.append({"vector": v["index"], "rank": rank, "rep": rep, "s": round(dt, 3), "pbs": p1 - p0, "levels": l1 - l0,
                          "signature_matches_reference": ok, "schedule": "fused" if fused else "faithful"})
    ctx.sync()
    t_all = time.perf_counter() - t_all
    for ln in lines:
        print(json.dumps(ln), flush=True)
    ok_all = all(ln["signature_matches_reference"] for ln in lines)
    if dist is not None:
        import torch
        t = torch.tensor([t_all, 0.0 if ok_all else 1.0, float(sum(ln["pbs"] for ln in lines))], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_all, bad, pbs = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        bad, pbs = 0.0 if ok_all else 1.0, float(sum(ln["pbs"] for ln in lines))
    if rank == 0:
        nsig = len(golden) * reps
        print(json.dumps({"summary": "sign_fhe_with_k0 over %d BIP-340 rows" % len(golden), "n_gpus": world, "signatures": nsig,
                          "wall_s": round(t_all, 3), "s_per_signature": round(t_all / nsig, 3), "pbs_total": int(pbs),
                          "pbs_per_s": round(pbs / t_all), "all_signatures_match": bad == 0.0,
                          "schedule": "fused" if fused else "faithful", "preset": preset,
                          "keygen_s": round(t_keygen, 2), "key_upload_s": round(t_upload, 2)}), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
