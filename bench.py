#!/usr/bin/env python
"""bench.py — batched PBS microbench (BASELINE.json configs[1]) + the operator / signing latencies of configs[2..4].

Headline (the JSON line's metric / value / e2e / roofline): 4096 LWE blocks, PARAM_MESSAGE_2_CARRY_2, keyswitch +
programmable bootstrap per block, on N B200s of one node.  One "step" = one pass of the hot path (keyswitch -> blind
rotation -> sample extraction) over one batch of 4096 synthetic random big-LWE ciphertexts per GPU.  `value` = whole-job
PBS/s with inputs resident in HBM (weak scaling: independent batches, no collective); `e2e` = the same through the
host-buffer C-ABI call (pinned host buffers, H2D and D2H inside the timed region).

Also in the same line, all measured in this run:
  `variants`      the headline kernel at the reference's accumulator width (64 bit) and on the TUniform n = 887 flavour;
  `ops`           BASELINE.json's other metric parts on REAL ciphertexts (product client: keygen, encrypt, decrypt), each
                  checked by decryption: 256-bit mul / encrypted shift / div 5, 514-bit mod n, fused k + e*d, and
                  Schnorr::sign_fhe_with_k0 over the 8 BIP-340 signing rows (bytes compared with the reference's plaintext
                  twin).  Under --gpus N these run with PBS levels sharded over the N ranks (STRONG scaling: one operator,
                  N GPUs; exchange fused into the blind rotation over peer-mapped pools), next to the weak-scaling PBS/s;
  `cpu_baseline`  the oracle port of the reference's path on the host cores (N = 1, rank 0): PBS/s and the perf_test.rs
                  operator list (src/perf_test.rs:27-80) through the same radix circuits over the CPU oracle.

`--impl reference` times the CPU oracle port of the reference's keyswitch+PBS (tfhe-rs itself cannot be built here: no
cargo, crate not vendored) on all host cores.

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
_JSON_OUT = sys.stdout
N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141        # src/scalar.rs:8
FP64_NOMINAL_TF = 148 * 64 * 2 * 1.965e9 / 1e12      # 148 SMs x 64 FMA/clk x 2 x 1.965 GHz = 37.2


# DRAM bytes (read + write) of one 4096-block launch, from the committed ncu --set full captures (NOT measured in this run)
NCU_TRAFFIC = {"pbs_ring_kernel": 202.69e6, "pbs_stream_kernel": 233.47e6, "pbs_stream_tx_kernel": 261.73e6}
NCU_TRAFFIC_SOURCE = "profiles/r02s_ncu_key_metrics.json / r01c_ncu_key_metrics.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch; constant, not re-measured here)"


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


# ------------------------------------------------------------------------------------------------
# CPU legs (the oracle port; test infrastructure used here only as the reported baseline)
# ------------------------------------------------------------------------------------------------
def cpu_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1: set the count explicitly."""
    from oracle import orc
    n = orc.host_cores()
    orc.set_threads(n)
    return n


def cpu_operator_times(okeys, threads, wide=False):
    """The reference's operator list (src/perf_test.rs:27-80) on the CPU: the same radix circuits (csrc/radix.cpp) over the
    oracle as their device (tests/host/oracle_backend.cpp), real ciphertexts, checked by decryption."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_client import OracleClientKey
    from oracle_radix import OracleRadix
    dev = OracleRadix(okeys, threads)
    R = dev.radix
    ck = OracleClientKey(okeys, seed=99)
    enc, dec = (lambda v, n: ck.encrypt_blocks(v, n, R)), (lambda r: ck.decrypt(r, R))
    a, b, c = enc(1344, 16), enc(5, 16), enc(7, 4)
    x256, y256 = 0x1234567890ABCDEF << 190 | 0xFEDCBA9876543211, 0x0FEDCBA987654321 << 188 | 0x13579BDF02468ACF
    ops = [
        ("u32 add (perf_test.rs:28)", lambda: a + b, 1349),
        ("u32 mul (perf_test.rs:32)", lambda: a * b, 6720),
        ("u32 shr by encrypted amount (perf_test.rs:36)", lambda: a >> b, 42),
        ("u8 min (perf_test.rs:44)", lambda: R.min(R.cast(a, 4), c), 7),
        ("u8 and 1 (perf_test.rs:48)", lambda: c & 1, 1),
        ("u32 div 5 (perf_test.rs:54)", lambda: a // 5, 268),
    ]
    if wide:
        A, B = enc(x256, 128), enc(y256, 128)
        ops.append(("256-bit add", lambda: A + B, (x256 + y256) % 2**256))
    out = []
    dec(c & 1)      # warm-up: FFT plan, LUT polynomials, thread pool
    for name, fn, want in ops:
        p0, l0 = R.stats()
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        p1, l1 = R.stats()
        out.append({"op": name, "ms": round(dt * 1e3, 1), "pbs": p1 - p0, "levels": l1 - l0, "correct": dec(r) == want})
    dev.close()
    return out


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's keyswitch+PBS on ALL host cores (rank 0 only; other ranks exit)."""
    if rank != 0:
        return
    from oracle import orc
    threads = cpu_threads()
    p = orc.preset(PRESET)
    K = orc.Keys(p, 1)
    sample = max(threads * 4, 16)
    cts = synthetic_batch(sample, 2049, 0xB200)
    lut = K.make_lut(np.arange(16))
    for _ in range(max(args.warmup, 1)):
        K.ks_pbs(cts[:threads], lut, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        K.ks_pbs(cts, lut, nthreads=threads)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "pbs_per_s", "value": v, "unit": "PBS/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": sample,
                   "note": "a rate: each step is a bounded sample (%d blocks) of the 4096-block batch, same parameter set, same LUTs; "
                           "one host, all %d cores, whatever --gpus says (the CPU path does not use GPUs)" % (sample, threads)},
        "cpu_baseline": {"value": v, "unit": "PBS/s", "cores": threads, "kind": "port",
                         "sample": "%d random LWE blocks per step, %d steps, oracle C port (OpenMP, %d threads), not tfhe-rs" % (sample, args.steps, threads)},
        "e2e": {"value": v, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_ops:
        try:
            line["cpu_baseline"]["ops"] = cpu_operator_times(K, threads, wide=True)
        except Exception as e:      # the baseline's operator leg must never cost the headline
            line["cpu_baseline"]["ops_error"] = repr(e)
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# GPU legs
# ------------------------------------------------------------------------------------------------
def headline_variant(fsb, preset, acc_bits, local, steps, F_of):
    """Short run of the same microbench on another accumulator width / parameter flavour (device-resident, CUDA events)."""
    from fhe_sign_b200.capi import LWE_BIG, LWE_SMALL
    params = fsb.Params.preset(preset, acc_bits=acc_bits)
    n = params.lwe_dim
    krng = np.random.default_rng(0x5EED)
    ctx = fsb.Context(params, device=local)
    ctx.upload_keys(krng.integers(0, 2**64, n * 4 * 2048, dtype=np.uint64),
                    krng.integers(0, 2**64, 2048 * params.ks_level * (n + 1), dtype=np.uint64))
    luts = ctx.luts_from_tables(np.stack([np.arange(16), (np.arange(16) * 3) % 16]))
    lut_idx = (np.arange(BATCH) % 2).astype(np.uint32)
    din = ctx.lwe(LWE_BIG, BATCH).upload(synthetic_batch(BATCH, 2049, 0xB210))
    dout, dsmall = ctx.lwe(LWE_BIG, BATCH), ctx.lwe(LWE_SMALL, BATCH)
    ctx.ks_pbs(din, luts, lut_idx, dout)
    ctx.sync()
    ctx.timer_start()
    for _ in range(steps):
        ctx.ks_pbs(din, luts, lut_idx, dout)
    ms = ctx.timer_stop() / steps
    ctx.keyswitch(din, dsmall)
    ctx.timer_start()
    for _ in range(steps):
        ctx.pbs(dsmall, luts, lut_idx, dout)
    ms_pbs = ctx.timer_stop() / steps
    name = ctx.pbs_kernel_name()
    for a in (din, dout, dsmall):
        a.free()
    ctx.close()
    return {"preset": preset, "lwe_dim": n, "acc_bits": acc_bits, "pbs_per_s": BATCH / (ms * 1e-3), "ms_per_step": ms,
            "kernel": name, "kernel_ms_per_launch": ms_pbs, "achieved_tflops": BATCH * F_of(n) / (ms_pbs * 1e-3) / 1e12, "steps": steps}


def run_ops(fsb, local, rank, world, dist, torch):
    """BASELINE.json configs[2..4] on real ciphertexts, checked by decryption.  world > 1: PBS levels sharded (strong scaling)."""
    from fhe_sign_b200 import biguint as bg
    from fhe_sign_b200 import schnorr
    from fhe_sign_b200.biguint import BigUintFHE
    from fhe_sign_b200.client import generate_keys

    t0 = time.perf_counter()
    # test-style reproducible keys AND encryption randomness: every rank must hold identical ciphertexts (SPMD contract)
    ck, (bsk, ksk) = generate_keys(PRESET, seed=2024, encryption_seed=7)
    t_keygen = time.perf_counter() - t0
    stream = torch.cuda.Stream() if world > 1 else None      # the NCCL-callback fallback enqueues its all-gather on the context's stream
    ctx = fsb.Context(fsb.Params.preset(PRESET, acc_bits=ACC_BITS), device=local, stream=stream.cuda_stream if stream else 0)
    t0 = time.perf_counter()
    ctx.upload_keys(bsk, ksk)
    t_upload = time.perf_counter() - t0
    R = ctx.radix
    exchange = "none (1 GPU)"
    peer_on = False
    if world > 1:
        from fhe_sign_b200.distributed import enable_level_sharding, enable_peer_sharding
        mode = os.environ.get("FSC_BENCH_EXCHANGE", "peer")
        min_width = int(os.environ.get("FSC_SHARD_MIN", "149"))
        ok = 0.0
        if mode == "peer":
            try:
                enable_peer_sharding(ctx, min_width=min_width)
                ok = 1.0
            except Exception as e:      # no peer access / IPC on this box: every rank must take the same path
                print("peer-mapped exchange unavailable on rank %d: %r" % (rank, e), file=sys.stderr)
        t = torch.tensor([ok], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if float(t[0]) == 1.0:
            peer_on = True
            exchange = "peer-mapped block pools: blind-rotation epilogue stores into every rank's pool over NVLink + flag barrier kernel (fsc_peer_pool_*)"
        else:
            if ok:
                from fhe_sign_b200.distributed import disable_peer_sharding
                disable_peer_sharding(ctx)
            enable_level_sharding(ctx, stream, min_width=min_width, capacity_blocks=1 << 16)
            exchange = "NCCL all-gather per level through the fsc_set_level_exchange callback (peer-mapped pools unavailable or FSC_BENCH_EXCHANGE=nccl)"

    def sync_all():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def tmax(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    rnd = np.random.default_rng(5)
    x, y, z = (int.from_bytes(rnd.bytes(32), "little") for _ in range(3))
    enc = lambda v, nb: ck.encrypt_blocks(v, nb, R)
    a, b, c = enc(x, 128), enc(y, 128), enc(z, 128)
    w = enc(x * y + z, 257)
    cases = [
        ("256-bit mul (128 blocks, wrapping)", "configs[2]", lambda: a * b, (x * y) % 2**256),
        ("256-bit shr by encrypted amount", "configs[2]", lambda: a >> b, x >> (y % 256)),
        ("256-bit add", "configs[0]", lambda: a + b, (x + y) % 2**256),
        ("256-bit div 5 (src/perf_test.rs:54 widened)", "configs[3]", lambda: a // 5, x // 5),
        ("514-bit mod n (secp256k1 order, src/scalar.rs:8)", "configs[3]", lambda: w % N_ORDER, (x * y + z) % N_ORDER),
        ("k + e*d fused (src/schnorr.rs:274), 256x256+256 -> 514 bits", "configs[4]", lambda: R.mul_add_wide(a, b, c, 272), x * y + z),
    ]
    ops, launches0 = [], ctx.launch_count()
    for name, cfg, fn, want in cases:
        best, dev_ms = None, None
        for rep in range(3):      # first repetition warms up (LUT uploads, pool growth); best of the three
            sync_all()
            p0, l0 = R.stats(); s0 = R.sharded_levels()
            ctx.timer_start()
            t0 = time.perf_counter()
            out = fn()
            d_ms = ctx.timer_stop()      # synchronises the context's stream
            dt = (time.perf_counter() - t0) * 1e3
            p1, l1 = R.stats(); s1 = R.sharded_levels()
            if best is None or dt < best:
                best, dev_ms = dt, d_ms
        ok = ck.decrypt(out, R) == want
        ops.append({"op": name, "config": cfg, "ms": round(tmax(best), 3), "device_ms": round(tmax(dev_ms), 3), "pbs": p1 - p0,
                    "levels": l1 - l0, "sharded_levels": s1 - s0, "correct": bool(ok)})
        del out

    # ---- Schnorr::sign_fhe_with_k0 over the BIP-340 signing rows (tests/golden/schnorr_vectors.json) --------------------
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "schnorr_vectors.json")))
    bg.set_server_key(ctx)

    def sign_row(v, public_challenge=False):
        d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
        p0, _ = R.stats()
        t0 = time.perf_counter()
        sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck, fused=True, public_challenge=public_challenge)
        dt = time.perf_counter() - t0
        p1, _ = R.stats()
        return dt, p1 - p0, sig.to_bytes().hex().upper() == v["reference_signature"]

    sign = {"schedule": "fused (fsc_radix_mul_add_wide), client-side encryption of e, k, d and decryption of s inside the timed region"}
    sign_row(golden[1])      # warm-up (LUTs, pool)
    sync_all()
    # (1) latency of ONE signature; world > 1: its PBS levels sharded over all ranks (every rank runs the same signature)
    dt, pbs, ok1 = sign_row(golden[1])
    sign["one_signature_s"] = round(tmax(dt), 4)
    sign["one_signature_pbs"] = pbs
    sign["one_signature_gpus"] = world
    all_ok = ok1
    # (1b) protocol-level variant, reported separately (NOT the reference's dataflow, which encrypts e too): the challenge e is
    # public by construction, so k + e*d can be a scalar x ciphertext product; d and k stay encrypted, same signature bytes
    sign_row(golden[1], public_challenge=True)
    sync_all()
    dt, pbs, okp = sign_row(golden[1], public_challenge=True)
    sign["public_challenge_variant"] = {"one_signature_s": round(tmax(dt), 4), "pbs": pbs, "match": bool(okp),
                                        "note": "e = H(R || P || m) in plaintext (every verifier recomputes it), d and k encrypted: "
                                                "fsc_radix_scalar_mul_add_wide; not the reference's dataflow"}
    all_ok = all_ok and okp
    # (2) all 8 rows: world == 1 sequentially; world > 1 as independent signatures, row r on rank r mod world, no exchange
    if world > 1:
        if peer_on:
            from fhe_sign_b200.distributed import disable_peer_sharding
            disable_peer_sharding(ctx)
        else:
            ctx.sync()
            dist.barrier()
            ctx._check(ctx.L.fsc_set_level_exchange(ctx.h, 0, 1, 0, None, 0, None, None))      # world = 1: sharding off
    sync_all()
    t0 = time.perf_counter()
    rows = []
    for i, v in enumerate(golden):
        if i % world != rank:
            continue
        dt, pbs, ok = sign_row(v)
        rows.append({"vector": v["index"], "s": round(dt, 4), "pbs": pbs, "match": bool(ok)})
        all_ok = all_ok and ok
    ctx.sync()
    t_rows = tmax(time.perf_counter() - t0)
    bad = tmax(0.0 if all_ok else 1.0)
    sign["all_8_rows_s"] = round(t_rows, 4)
    sign["rows_per_gpu"] = "row r on rank r mod %d (independent signatures, no exchange)" % world if world > 1 else "sequential on one GPU"
    sign["rank0_rows"] = rows
    sign["all_signatures_match"] = bad == 0.0
    launches = ctx.launch_count() - launches0
    ctx.close()
    return {"scaling": "strong" if world > 1 else "1 GPU", "n_gpus": world, "exchange": exchange, "preset": PRESET, "acc_bits": ACC_BITS,
            "keygen_s": round(t_keygen, 2), "key_upload_s": round(t_upload, 2),
            "timing": "ms = wall clock around the operator call + stream sync (host scheduling included), max over ranks, best of 3; "
                      "device_ms = CUDA events on the context's stream",
            "operators": ops, "all_correct": all(o["correct"] for o in ops), "sign_fhe_with_k0": sign, "gpu_launches": int(launches)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ops", action="store_true", help="headline only (skip the operator / signing block and the variants)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # ONE JSON line on stdout: everything else any library prints there (NCCL's version banner, for one) goes to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

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
        narrow = {"note": "independent blocks, at most one per SM: what a carry-propagation level of a radix operator costs", "levels": []}
        for nb in (16, 128):
            dsm_n, dout_n = ctx.lwe(LWE_SMALL, nb), ctx.lwe(LWE_BIG, nb)
            dsm_n.upload(np.ascontiguousarray(synthetic_batch(nb, n + 1, 0xB201)))
            ctx.pbs(dsm_n, luts, None, dout_n)
            ctx.sync()
            ctx.timer_start()
            for _ in range(3):
                ctx.pbs(dsm_n, luts, None, dout_n)
            narrow["levels"].append({"blocks": nb, "ms_per_level": ctx.timer_stop() / 3})
            dsm_n.free(); dout_n.free()
        narrow["blocks"], narrow["ms_per_level"] = 128, narrow["levels"][-1]["ms_per_level"]

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
    kernel_name = ctx.pbs_kernel_name()
    for arr in (din, dout, dsmall):
        arr.free()
    ctx.close()
    del host_in, host_out

    # ---- the same kernel at the reference's accumulator width and on the other parameter flavour (rank 0, N = 1 only) ----
    variants = None
    if rank == 0 and world == 1 and not args.no_ops:
        variants = []
        other_preset = "2_2_tuniform" if PRESET == "2_2_gaussian" else "2_2_gaussian"
        for pre, acc in ((PRESET, 96 - ACC_BITS), (other_preset, ACC_BITS), (other_preset, 96 - ACC_BITS)):
            try:
                v = headline_variant(fsb, pre, acc, local, max(2, min(args.steps, 3)), flops_per_pbs)
                v["frac"] = v["achieved_tflops"] / fp64_peak if fp64_peak else None
                v["frac_vs_nominal"] = v["achieved_tflops"] / FP64_NOMINAL_TF
                variants.append(v)
            except Exception as e:
                variants.append({"preset": pre, "acc_bits": acc, "error": repr(e)})

    # ---- operators and signing on real ciphertexts (every rank takes part: SPMD) ------------------------------------------
    ops = None
    if not args.no_ops:
        try:
            ops = run_ops(fsb, local, rank, world, dist, torch)
        except Exception as e:
            if world > 1:
                raise      # a rank that drops out would leave its peers in a barrier
            ops = {"error": repr(e)}

    if rank == 0:
        peaks = measured_peaks()
        ms_step = ms_total / args.steps
        value = world * BATCH * args.steps / (ms_total * 1e-3)
        F = flops_per_pbs(n)
        achieved_tf = BATCH * F / (ms_pbs * 1e-3) / 1e12
        hbm_bytes = BATCH * (n + 1) * 8 + BATCH * words * 8 + n * 65536     # small LWE in, big LWE out, Fourier BSK once
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = NCU_TRAFFIC.get(kernel_name) if (BATCH == 4096 and n == 834 and ACC_BITS == 32) else None
        line = {
            "metric": "pbs_per_s", "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "lwe_dim": n, "acc_bits": ACC_BITS,
                       "l2": "per-step working set 2x67 MB ciphertexts + 123 MB keys > 126 MB L2 (inputs larger than L2)",
                       "ops_block": "BASELINE configs[2..4] measured in the same run under `ops` (strong scaling under --gpus N)"},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None,
                         "peak_nominal": FP64_NOMINAL_TF, "frac_vs_nominal": achieved_tf / FP64_NOMINAL_TF,
                         "traffic": traffic, "traffic_source": NCU_TRAFFIC_SOURCE if traffic else None,
                         "kernel": kernel_name, "ms_per_launch": ms_pbs, "flops_per_pbs": F,
                         "peak_source": "measured in this run (fsc_measure_fp64_peak: 16 FMA chains per thread, two fresh operand pairs per DFMA, 512 FMAs per loop "
                                        "trip, 4 warps per SM sub-partition - the highest-reading pattern of tools/ubench/fp64_operands.cu; MEASURED_PEAKS.json has "
                                        "no FP64 figure); peak_nominal = 148 SM x 64 FMA/clk x 2 x 1.965 GHz",
                         "keyswitch_ms_per_launch": ms_ks,
                         "hbm": {"achieved": hbm_bytes / (ms_pbs * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_bytes / (ms_pbs * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}},
            "e2e": {"value": world * BATCH * args.steps / (e2e_ms * 1e-3), "unit": "PBS/s",
                    "h2d_bytes_per_step": int(BATCH * words * 8 + BATCH * 4), "d2h_bytes_per_step": int(BATCH * words * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "narrow_level": narrow,
            "variants": variants,
            "ops": ops,
        }
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
            from oracle import orc
            threads = cpu_threads()
            okeys = orc.Keys(orc.preset(PRESET), 1)
            sample = max(threads * 8, 32)
            cts = synthetic_batch(sample, words, 0xB200)
            lut = okeys.make_lut(np.arange(16))
            okeys.ks_pbs(cts[:threads], lut, nthreads=threads)
            t0 = time.perf_counter()
            okeys.ks_pbs(cts, lut, nthreads=threads)
            dt = time.perf_counter() - t0
            cpu = {"value": sample / dt, "unit": "PBS/s", "cores": threads, "kind": "port",
                   "sample": "first %d blocks of the same batch, oracle C port (OpenMP, %d threads), not tfhe-rs" % (sample, threads)}
            if not args.no_ops:
                try:
                    cpu["ops"] = cpu_operator_times(okeys, threads, wide=False)
                    if ops and "operators" in ops:
                        kd = [o for o in ops["operators"] if o["op"].startswith("k + e*d")]
                        if kd:
                            cpu["k_plus_ed_extrapolated_s"] = round(kd[0]["pbs"] / cpu["value"], 1)
                            cpu["k_plus_ed_note"] = ("EXTRAPOLATED: the fused schedule's %d bootstraps / the measured CPU PBS rate (wide levels keep all cores "
                                                     "busy); a timed CPU run of the same circuit is kept in profiles/" % kd[0]["pbs"])
                except Exception as e:
                    cpu["ops_error"] = repr(e)
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
