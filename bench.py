#!/usr/bin/env python
"""bench.py — batched PBS microbench (BASELINE.json configs[1]): 4096 LWE blocks,
PARAM_MESSAGE_2_CARRY_2, keyswitch + programmable bootstrap per block, on N B200s of one node.

One "step" = one pass of the hot path (keyswitch -> blind rotation -> sample extraction) over one
batch of 4096 synthetic random big-LWE ciphertexts per GPU.  `value` = whole-job PBS/s with inputs
resident in HBM; `e2e` = the same through the host-buffer C-ABI call (pinned host buffers, H2D and
D2H inside the timed region).  `--impl reference` times the CPU oracle port of the reference's path
(tfhe-rs itself cannot be built here: no cargo, crate not vendored) on the host cores.

Launch: python bench.py [--gpus N --steps K --warmup W]   (N > 1: under torchrun, one rank per GPU)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESET = os.environ.get("FSC_BENCH_PRESET", "2_2_gaussian")
BATCH = int(os.environ.get("FSC_BENCH_BATCH", "4096"))
ACC_BITS = int(os.environ.get("FSC_BENCH_ACC_BITS", "32"))
WORKLOAD = "batched PBS microbench: %d LWE blocks, PARAM_MESSAGE_2_CARRY_2 (%s), keyswitch+PBS per block" % (BATCH, PRESET)


# DRAM bytes (read + write) of one 4096-block launch, from the ncu --set full captures summarised in profiles/
NCU_TRAFFIC = {"pbs_ring_kernel": 202.69e6, "pbs_stream_kernel": 233.47e6}      # profiles/r01c_ncu_key_metrics.json


def flops_per_pbs(n, N=2048, k=1, l=1):
    """n * [((k+1)l + (k+1)) * (5 M log2 M + 6 M) + 8 (k+1)^2 l M], M = N/2."""
    M = N // 2
    lg = M.bit_length() - 1
    return n * (((k + 1) * l + (k + 1)) * (5 * M * lg + 6 * M) + 8 * (k + 1) ** 2 * l * M)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(int(f[0])); mx.append(int(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(count, words, seed):
    """uniformly random ciphertext words (PBS cost is data independent)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 2**64, (count, words), dtype=np.uint64)


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's keyswitch+PBS on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import orc
    p = orc.preset(PRESET)
    K = orc.Keys(p, 1)
    threads = orc.max_threads()
    sample = max(threads * 4, 16)
    cts = synthetic_batch(sample, 2049, 0xB200)
    lut = K.make_lut(np.arange(16))
    for _ in range(max(args.warmup, 1)):
        K.ks_pbs(cts[:threads], lut)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        K.ks_pbs(cts, lut)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "pbs_per_s", "value": v, "unit": "PBS/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": sample, "note": "bounded sample of the 4096-block batch"},
        "cpu_baseline": {"value": v, "unit": "PBS/s", "cores": threads, "kind": "port",
                         "sample": "%d random LWE blocks per step, %d steps, oracle C port (OpenMP), not tfhe-rs" % (sample, args.steps)},
        "e2e": {"value": v, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import fhe_sign_b200 as fsb
    from fhe_sign_b200.capi import LWE_BIG

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # synthetic key material of the right shape (PBS/keyswitch cost is data independent); the oracle is
    # not touched by this arm except in the cpu_baseline leg below
    params = fsb.Params.preset(PRESET, acc_bits=ACC_BITS)
    n = params.lwe_dim
    words = 2049
    krng = np.random.default_rng(0x5EED)
    ctx = fsb.Context(params, device=local)
    ctx.upload_keys(krng.integers(0, 2**64, n * 4 * 2048, dtype=np.uint64),
                    krng.integers(0, 2**64, 2048 * params.ks_level * (n + 1), dtype=np.uint64))

    host_in = torch.from_numpy(synthetic_batch(BATCH, words, 0xB200 + rank).view(np.int64)).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    np_in, np_out = host_in.numpy().view(np.uint64), host_out.numpy().view(np.uint64)
    luts = ctx.luts_from_tables(np.stack([np.arange(16), (np.arange(16) * 3) % 16]))
    lut_idx = (np.arange(BATCH) % 2).astype(np.uint32)
    din, dout = ctx.lwe(LWE_BIG, BATCH).upload(np_in), ctx.lwe(LWE_BIG, BATCH)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    fp64_peak = ctx.measure_fp64_peak()

    # ---- device-resident steps -------------------------------------------------------------
    for _ in range(args.warmup):
        ctx.ks_pbs(din, luts, lut_idx, dout)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.ks_pbs(din, luts, lut_idx, dout)
    ms_total = ctx.timer_stop()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    barrier()

    # per-kernel split (same stream, CUDA events), used for the roofline of the dominant kernel
    from fhe_sign_b200.capi import LWE_SMALL
    dsmall = ctx.lwe(LWE_SMALL, BATCH)
    ctx.keyswitch(din, dsmall)
    ctx.sync()
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.keyswitch(din, dsmall)
    ms_ks = ctx.timer_stop() / args.steps
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.pbs(dsmall, luts, lut_idx, dout)
    ms_pbs = ctx.timer_stop() / args.steps
    barrier()

    # latency of one narrow PBS level (at most one ciphertext per SM: the carry-propagation levels of every radix operator)
    narrow = None
    if rank == 0:
        nb = 128
        dsm_n, dout_n = ctx.lwe(LWE_SMALL, nb), ctx.lwe(LWE_BIG, nb)
        dsm_n.upload(np.ascontiguousarray(synthetic_batch(nb, n + 1, 0xB201)))
        ctx.pbs(dsm_n, luts, None, dout_n)
        ctx.sync()
        ctx.timer_start()
        for _ in range(3):
            ctx.pbs(dsm_n, luts, None, dout_n)
        narrow = {"blocks": nb, "ms_per_level": ctx.timer_stop() / 3,
                  "note": "128 independent blocks, one per SM: what a carry-propagation level of a 256-bit operator costs"}
        dsm_n.free(); dout_n.free()

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------
    for _ in range(max(1, min(args.warmup, 2))):
        ctx.apply_lut_host(np_in, luts, lut_idx, out=np_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.apply_lut_host(np_in, luts, lut_idx, out=np_out)
    ctx.sync()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        peaks = measured_peaks()
        ms_step = ms_total / args.steps
        value = world * BATCH * args.steps / (ms_total * 1e-3)
        F = flops_per_pbs(n)
        kernel_name = ctx.pbs_kernel_name()
        achieved_tf = BATCH * F / (ms_pbs * 1e-3) / 1e12
        hbm_bytes = BATCH * (n + 1) * 8 + BATCH * words * 8 + n * 65536     # small LWE in, big LWE out, Fourier BSK once
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": "pbs_per_s", "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "lwe_dim": n, "acc_bits": ACC_BITS,
                       "l2": "per-step working set 2x67 MB ciphertexts + 123 MB keys > 126 MB L2 (inputs larger than L2)"},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one 4096-block launch, ncu --set full capture
                         # (profiles/r01_ncu_key_metrics.json, n=834, acc 32)
                         "traffic": NCU_TRAFFIC.get(kernel_name) if (BATCH == 4096 and n == 834 and ACC_BITS == 32) else None,
                         "kernel": kernel_name, "ms_per_launch": ms_pbs, "flops_per_pbs": F,
                         "peak_source": "measured in this run (fsc_measure_fp64_peak; MEASURED_PEAKS.json has no FP64 figure)",
                         "keyswitch_ms_per_launch": ms_ks,
                         "hbm": {"achieved": hbm_bytes / (ms_pbs * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_bytes / (ms_pbs * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}},
            "e2e": {"value": world * BATCH * args.steps / (e2e_ms * 1e-3), "unit": "PBS/s",
                    "h2d_bytes_per_step": int(BATCH * words * 8 + BATCH * 4), "d2h_bytes_per_step": int(BATCH * words * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "narrow_level": narrow,
        }
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only (torchrun pins OMP_NUM_THREADS=1)
            from oracle import orc
            okeys = orc.Keys(orc.preset(PRESET), 1)
            threads = orc.max_threads()
            sample = max(threads * 8, 32)
            lut = okeys.make_lut(np.arange(16))
            okeys.ks_pbs(np_in[:threads], lut)
            t0 = time.perf_counter()
            okeys.ks_pbs(np_in[:sample], lut)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": "PBS/s", "cores": threads, "kind": "port",
                                    "sample": "first %d blocks of the same batch, oracle C port (OpenMP), not tfhe-rs" % sample}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
