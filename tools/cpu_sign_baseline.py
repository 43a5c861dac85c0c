"""The same-host CPU number behind "full FHE Schnorr signing >= 100x faster than the CPU reference" (BASELINE.json north_star),
TIMED rather than extrapolated: k + e*d for BIP-340 vector 1 (src/schnorr.rs:274) through the SAME radix circuits the GPU runs
(csrc/radix.cpp, fused schedule: 58 182 bootstraps in 18 levels) with the CPU oracle as the device (tests/host/oracle_backend.cpp:
keyswitch + PBS per level on all host cores, OpenMP), real ciphertexts, result decrypted and compared.
It is the oracle port, not tfhe-rs (which cannot be built here): say so wherever the ratio is quoted.

    python tools/cpu_sign_baseline.py            # about 2-3 minutes on 16-32 cores; prints one JSON line
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fhe_sign_b200 import biguint as bg  # noqa: E402
from fhe_sign_b200 import schnorr  # noqa: E402
from fhe_sign_b200.biguint import BigUintFHE  # noqa: E402
from oracle import orc  # noqa: E402
from oracle_client import OracleClientKey  # noqa: E402
from oracle_radix import OracleRadix  # noqa: E402


def main():
    threads = orc.host_cores()
    orc.set_threads(threads)
    K = orc.Keys(orc.preset("2_2_gaussian"), 1)
    dev = OracleRadix(K, threads)
    bg.set_server_key(dev)
    ck = OracleClientKey(K, seed=5)
    v = json.load(open(os.path.join(ROOT, "tests", "golden", "schnorr_vectors.json")))[1]
    d, k0, msg = int(v["secret_key"], 16), int(v["k0"], 16), bytes.fromhex(v["message"])
    p0, l0 = dev.radix.stats()
    t0 = time.perf_counter()
    sig = schnorr.sign_fhe_with_k0(msg, k0, d, BigUintFHE.new(d, ck), ck, fused=True)
    dt = time.perf_counter() - t0
    p1, l1 = dev.radix.stats()
    ok = sig.to_bytes().hex().upper() == v["reference_signature"]
    print(json.dumps({"what": "sign_fhe_with_k0, BIP-340 vector 1, fused schedule, on the CPU oracle (not tfhe-rs)", "seconds": round(dt, 1),
                      "cores": threads, "pbs": p1 - p0, "levels": l1 - l0, "pbs_per_s": round((p1 - p0) / dt, 1),
                      "signature_matches_reference": ok}), flush=True)
    dev.close()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
