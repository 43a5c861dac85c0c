"""Driver entry points: build() compiles every native piece for sm_100a; smoke() runs one small
invocation of the hot path on cuda:0 and checks it against the CPU oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build():
    # CUDA library (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo; cross-compiles without a GPU)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "fhe_sign_b200", "csrc"), "-s"])
    # the checker: CPU oracle and the lane-by-lane emulation harness (test infrastructure)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-shared", "-fPIC",
                           "-o", os.path.join(emu, "libpbs_emu.so"), os.path.join(emu, "pbs_emu.cpp")])
    import fhe_sign_b200
    fhe_sign_b200.load_library()


def smoke():
    import numpy as np
    import fhe_sign_b200 as fsb
    from oracle import orc

    K = orc.Keys(orc.preset("toy"), 1)
    for acc_bits in (64, 32):
        ctx = fsb.Context(fsb.Params.preset("toy", acc_bits=acc_bits), device=0)
        ctx.upload_keys(K.bsk, K.ksk)
        table = (np.arange(16) * 5 + 2) % 16
        luts = ctx.luts_from_tables(table)
        m = np.arange(32).astype(np.uint64) % 16
        ct = K.encrypt_msgs(m)
        out = ctx.apply_lut_host(ct, luts)
        got, ref = K.decrypt_msgs(out), K.decrypt_msgs(K.ks_pbs(ct, K.make_lut(table)))
        assert (got == table[m]).all() and (ref == table[m]).all(), (got, ref)
        ctx.close()
    print("smoke ok: GPU keyswitch+PBS decrypts identically to the CPU oracle (acc 64 and 32)")


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
