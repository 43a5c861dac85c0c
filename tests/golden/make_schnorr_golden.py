"""Generates tests/golden/schnorr_vectors.json from the reference's own fixture
(/root/reference/tests/test_vectors.csv, the official BIP-340 vectors compiled into
src/schnorr.rs:532) by replaying the reference's signing dataflow (oracle/schnorr_model.py).
Run in the build container only; the GPU box reads the committed JSON."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import schnorr_model as sm  # noqa: E402

rows = list(csv.DictReader(open("/root/reference/tests/test_vectors.csv")))
out = []
for r in rows:
    if not r["secret key"]:
        continue
    d = int(r["secret key"], 16)
    msg, aux = bytes.fromhex(r["message"]), bytes.fromhex(r["aux_rand"])
    pub = sm.pubkey_even_y(d)
    k0 = sm.compute_nonce(d, pub, msg, aux)              # src/schnorr.rs:81 (un-negated key, as the reference does)
    R, k, e = sm.signing_inputs(msg, k0, d)
    sig = sm.sign_with_k0(msg, k0, d)
    sig_fhe, s_wo = sm.sign_fhe_with_k0_model(msg, k0, d)
    assert s_wo == k + e * d, "vector triggers the dropped-carry path of biguint.rs Mul"
    out.append({
        "index": int(r["index"]), "secret_key": "%064X" % d, "public_key": r["public key"], "aux_rand": r["aux_rand"],
        "message": r["message"], "csv_signature": r["signature"], "k0": "%064X" % k0, "R_y_odd": bool(R[1] % 2),
        "e": "%064X" % e, "k": "%064X" % k, "k_plus_e_d": "%X" % s_wo, "reference_signature": sig.hex().upper(),
        "reference_matches_csv": sig.hex().upper() == r["signature"].upper(),
        "len_e": len(sm.to_u32_digits(e)), "len_d": len(sm.to_u32_digits(d)), "len_k": len(sm.to_u32_digits(k)),
    })
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "schnorr_vectors.json"), "w"), indent=1)
print("wrote", len(out), "vectors;", sum(v["reference_matches_csv"] for v in out), "match the CSV signature")
