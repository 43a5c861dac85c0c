"""Parity tests of the CUDA hot path against the CPU oracle, through the C ABI (run with -m gpu).

Keyswitch is integer work: bit-exact.  PBS outputs are not bit-comparable between two f64 FFT
implementations (rounding differences are re-randomised by the gadget decomposition of the next CMUX
step), so the bar there is the one the reference's own tests use — identical decrypted values —
plus noise statistics against the oracle's and against the parameter set's budget."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _noise(K, ct, expect_msgs):
    return (K.phase_big(ct) - K.encode(expect_msgs)).astype(np.int64).astype(np.float64) / 2.0**64


def test_fft_hook_matches_exact_negacyclic_product(gpu_ctx, orc, rng):
    ctx = gpu_ctx("toy")
    a = rng.integers(0, 2**64, (5, 2048), dtype=np.uint64)
    b = rng.integers(-2**22, 2**22, (5, 2048), dtype=np.int64)
    b[0] = 0; b[0, 0] = 1                                  # identity
    b[1] = 0; b[1, 2047] = 1                               # X^(N-1): exercises the negacyclic wrap
    c = ctx.debug_negacyclic_mul(a, b)
    for i in range(5):
        d = (c[i] - orc.negacyclic_mul_exact(a[i], b[i])).astype(np.int64)
        assert np.abs(d).max() < 2**43, i


@pytest.mark.parametrize("preset,count", [("toy", 1), ("toy", 70), ("2_2_gaussian", 33), ("2_2_tuniform", 16)])
def test_keyswitch_bit_exact(gpu_ctx, oracle_keys, rng, preset, count):
    from fhe_sign_b200.capi import LWE_BIG, LWE_SMALL
    K, ctx = oracle_keys(preset), gpu_ctx(preset)
    ct = K.encrypt_msgs(rng.integers(0, 16, count).astype(np.uint64))
    ct[0, :8] = [0, 2**64 - 1, 2**63, 2**63 - 1, 2**48, 2**48 - 1, 3 << 47, 1 << 60]   # decomposition edge values
    din, dout = ctx.lwe(LWE_BIG, count).upload(ct), ctx.lwe(LWE_SMALL, count)
    ctx.keyswitch(din, dout)
    got = dout.download()
    assert np.array_equal(got, K.keyswitch(ct))
    din.free(); dout.free()


@pytest.mark.parametrize("variant", ["simt", "mma", "umma"])
def test_keyswitch_variants_bit_exact(oracle_keys, rng, variant, monkeypatch):
    """the three keyswitch kernels (CUDA cores, mma.sync int8 limb GEMM, tcgen05 int8 limb GEMM) against the oracle, ragged counts."""
    import fhe_sign_b200 as fsb
    from fhe_sign_b200.capi import LWE_BIG, LWE_SMALL
    monkeypatch.setenv("FSC_KS_VARIANT", variant)
    K = oracle_keys("2_2_gaussian")
    ctx = fsb.Context(fsb.Params.preset("2_2_gaussian"))
    ctx.upload_keys(K.bsk, K.ksk)
    for count in (1, 129, 300):
        ct = rng.integers(0, 2**64, (count, 2049), dtype=np.uint64)        # arbitrary words: keyswitch is linear integer work
        ct[0, :6] = [0, 2**64 - 1, 2**63, 2**48, 2**48 - 1, 3 << 47]
        din, dout = ctx.lwe(LWE_BIG, count).upload(ct), ctx.lwe(LWE_SMALL, count)
        ctx.keyswitch(din, dout)
        assert np.array_equal(dout.download(), K.keyswitch(ct)), (variant, count)
        din.free(); dout.free()
    ctx.close()


@pytest.mark.parametrize("acc_bits", [64, 32])
@pytest.mark.parametrize("preset", ["toy", "2_2_gaussian", "2_2_tuniform"])
def test_pbs_all_messages_all_luts(gpu_ctx, oracle_keys, rng, preset, acc_bits):
    from fhe_sign_b200.capi import LWE_BIG
    K, ctx = oracle_keys(preset), gpu_ctx(preset, acc_bits)
    tables = np.stack([np.arange(16), (np.arange(16) * 3 + 1) % 16, rng.integers(0, 16, 16),
                       [((i >> 2) * (i & 3)) % 4 for i in range(16)]]).astype(np.uint64)
    luts = ctx.luts_from_tables(tables)
    m = np.tile(np.arange(16), 4).astype(np.uint64)
    idx = np.repeat(np.arange(4), 16).astype(np.uint32)
    ct = K.encrypt_msgs(m)
    din, dout = ctx.lwe(LWE_BIG, m.size).upload(ct), ctx.lwe(LWE_BIG, m.size)
    ctx.ks_pbs(din, luts, idx, dout)
    out = dout.download()
    assert (K.decrypt_msgs(out) == tables[idx, m]).all()
    assert np.abs(_noise(K, out, tables[idx, m])).max() < 2.0**-7
    din.free(); dout.free(); luts.free()


@pytest.mark.parametrize("acc_bits", [32, 64])
@pytest.mark.parametrize("variant", ["auto", "stream", "ring", "pair", "split", "solo", "quad", "duo", "stream-tx0", "stream-tx1", "auto-ring"])
def test_pbs_kernel_variants_all_widths(oracle_keys, orc, rng, variant, acc_bits, monkeypatch):
    """Every blind-rotation kernel (FSC_PBS_VARIANT) at batch widths that select each of its configurations
    (1, 2 and 3-4 ciphertexts per CTA, ragged last CTA): decrypted values equal the table, noise inside the budget.
    The Fourier key layout follows the variant, so this also checks both key conversions."""
    import fhe_sign_b200 as fsb
    from fhe_sign_b200.capi import LWE_BIG
    if variant in ("solo", "quad", "duo", "stream-tx0", "stream-tx1") and acc_bits != 32:
        pytest.skip("the solo, quad, duo and tensor-memory stream kernels exist for the 32-bit accumulator only")
    if variant.startswith("stream-tx"):      # comparison forms of the wide-batch stream kernel (default: FSC_STREAM_TX=2)
        monkeypatch.setenv("FSC_STREAM_TX", variant[-1])
        variant = "stream"
    if variant == "auto-ring":               # the ring kernel for wide batches (the default before pbs_stream_tx_kernel<2>)
        monkeypatch.setenv("FSC_PBS_WIDE", "ring")
        variant = "auto"
    monkeypatch.setenv("FSC_PBS_VARIANT", variant)
    K = oracle_keys("toy")
    ctx = fsb.Context(fsb.Params.preset("toy", acc_bits=acc_bits))
    ctx.upload_keys(K.bsk, K.ksk)
    table = rng.integers(0, 16, 16).astype(np.uint64)
    luts = ctx.luts_from_tables(table)
    sms = 148
    for count in (3, sms + 5, 2 * sms + 7):
        m = rng.integers(0, 16, count).astype(np.uint64)
        din, dout = ctx.lwe(LWE_BIG, count).upload(K.encrypt_msgs(m)), ctx.lwe(LWE_BIG, count)
        ctx.ks_pbs(din, luts, None, dout)
        out = dout.download()
        assert (K.decrypt_msgs(out) == table[m]).all(), (variant, acc_bits, count)
        assert np.abs(_noise(K, out, table[m])).max() < 2.0**-7
        din.free(); dout.free()
    a = rng.integers(0, 2**64, (2, 2048), dtype=np.uint64)
    b = rng.integers(-2**22, 2**22, (2, 2048), dtype=np.int64)
    c = ctx.debug_negacyclic_mul(a, b)
    for i in range(2):
        assert np.abs((c[i] - orc.negacyclic_mul_exact(a[i], b[i])).astype(np.int64)).max() < 2**43
    ctx.close()


def test_lut_polynomials_match_oracle(gpu_ctx, oracle_keys, rng):
    """fsc_luts_from_tables builds the same accumulator the oracle does: PBS through uploaded oracle
    polynomials and through library-built ones decrypt identically."""
    from fhe_sign_b200.capi import LWE_BIG
    K, ctx = oracle_keys("toy"), gpu_ctx("toy")
    table = rng.integers(0, 16, 16).astype(np.uint64)
    l1, l2 = ctx.luts_from_tables(table), ctx.luts_upload(K.make_lut(table))
    m = np.arange(16).astype(np.uint64)
    din = ctx.lwe(LWE_BIG, 16).upload(K.encrypt_msgs(m))
    o1, o2 = ctx.lwe(LWE_BIG, 16), ctx.lwe(LWE_BIG, 16)
    ctx.ks_pbs(din, l1, None, o1)
    ctx.ks_pbs(din, l2, None, o2)
    assert np.array_equal(o1.download(), o2.download())      # same kernel, same inputs: deterministic
    assert (K.decrypt_msgs(o1.download()) == table[m]).all()


@pytest.mark.parametrize("acc_bits", [64, 32])
def test_pbs_noise_matches_oracle_toy(gpu_ctx, oracle_keys, rng, acc_bits):
    """Same inputs through the GPU and through the oracle: equal decryptions, equal noise level."""
    from fhe_sign_b200.capi import LWE_SMALL, LWE_BIG
    K, ctx = oracle_keys("toy"), gpu_ctx("toy", acc_bits)
    count = 512
    m = rng.integers(0, 16, count).astype(np.uint64)
    small = K.keyswitch(K.encrypt_msgs(m))
    luts = ctx.luts_from_tables(np.arange(16))
    din, dout = ctx.lwe(LWE_SMALL, count).upload(small), ctx.lwe(LWE_BIG, count)
    ctx.pbs(din, luts, None, dout)
    out = dout.download()
    ref = K.pbs(small, K.make_lut(np.arange(16)))
    assert (K.decrypt_msgs(out) == m).all() and (K.decrypt_msgs(ref) == m).all()
    s_gpu, s_ref = _noise(K, out, m).std(), _noise(K, ref, m).std()
    assert s_gpu < 1.15 * s_ref, (s_gpu, s_ref)


def test_apply_lut_host_ragged_and_empty(gpu_ctx, oracle_keys, rng):
    K, ctx = oracle_keys("toy"), gpu_ctx("toy")
    luts = ctx.luts_from_tables(np.stack([np.arange(16), 15 - np.arange(16)]))
    for count in (0, 1, 37, 1500):      # 1500: more than one and a half kernel waves, so uploads / downloads run chunked on the copy streams
        m = rng.integers(0, 16, count).astype(np.uint64)
        idx = rng.integers(0, 2, count).astype(np.uint32)
        out = ctx.apply_lut_host(K.encrypt_msgs(m).reshape(count, 2049), luts, idx)
        assert out.shape == (count, 2049)
        assert (K.decrypt_msgs(out) == np.where(idx == 0, m, 15 - m)).all()


def test_apply_lut_host_chunk_plans_agree_bit_for_bit(oracle_keys, rng, monkeypatch):
    """fsc_apply_lut_host cuts a batch of more than 2.5 kernel waves into first wave | middle | last wave and runs the blind
    rotations of consecutive chunks on two alternating streams; uniform chunks (FSC_HOST_CHUNK_WAVES) stay on one stream.  The
    kernels are deterministic per ciphertext, so every plan and the device-resident call must give identical ciphertexts."""
    import fhe_sign_b200 as fsb
    from fhe_sign_b200.capi import LWE_BIG
    K = oracle_keys("toy")
    count = 4 * 148 * 3 + 400                      # 3.7 waves: three chunks by default, each wide enough for the same (wide-batch) kernel
    m = rng.integers(0, 16, count).astype(np.uint64)
    cts = K.encrypt_msgs(m).reshape(count, 2049)
    idx = rng.integers(0, 2, count).astype(np.uint32)
    outs = []
    for waves in (None, "1", "0"):
        if waves is None:
            monkeypatch.delenv("FSC_HOST_CHUNK_WAVES", raising=False)
        else:
            monkeypatch.setenv("FSC_HOST_CHUNK_WAVES", waves)
        ctx = fsb.Context(fsb.Params.preset("toy", acc_bits=32))
        ctx.upload_keys(K.bsk, K.ksk)
        luts = ctx.luts_from_tables(np.stack([np.arange(16), 15 - np.arange(16)]))
        outs.append(ctx.apply_lut_host(cts, luts, idx).copy())
        if waves is None:
            din, dout = ctx.lwe(LWE_BIG, count).upload(cts), ctx.lwe(LWE_BIG, count)
            ctx.ks_pbs(din, luts, idx, dout)
            outs.append(dout.download().copy())
        ctx.close()
    assert (K.decrypt_msgs(outs[0]) == np.where(idx == 0, m, 15 - m)).all()
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])


def test_padding_bit_negates(gpu_ctx, oracle_keys):
    """Messages with the padding bit set (m >= 16) come out negated: the negacyclic property every
    radix circuit has to respect."""
    K, ctx = oracle_keys("toy"), gpu_ctx("toy")
    luts = ctx.luts_from_tables(np.arange(16))
    m = np.arange(16, 32).astype(np.uint64)
    out = ctx.apply_lut_host(K.encrypt_msgs(m), luts)
    assert (K.decrypt_msgs(out) == (32 - (m - 16)) % 32).all()


def test_error_codes(gpu_ctx, oracle_keys):
    import fhe_sign_b200 as fsb
    from fhe_sign_b200.capi import LWE_BIG, LWE_SMALL
    ctx = fsb.Context(fsb.Params.preset("toy"))
    luts = ctx.luts_from_tables(np.arange(16))
    a, b = ctx.lwe(LWE_BIG, 4), ctx.lwe(LWE_BIG, 4)
    with pytest.raises(fsb.FscError) as ei:
        ctx.ks_pbs(a, luts, None, b)
    assert ei.value.code == 5                                  # keys not uploaded
    K = oracle_keys("toy")
    with pytest.raises(fsb.FscError) as ei:
        ctx.upload_keys(K.bsk[:-1], K.ksk)
    assert ei.value.code == 1
    ctx.upload_keys(K.bsk, K.ksk)
    with pytest.raises(fsb.FscError) as ei:
        ctx.ks_pbs(a, luts, np.array([0, 1, 0, 0], dtype=np.uint32), b)    # lut index out of range
    assert ei.value.code == 1
    with pytest.raises(fsb.FscError) as ei:
        ctx.ks_pbs(a, luts, None, ctx.lwe(LWE_SMALL, 4))      # wrong kind
    assert ei.value.code == 1
    with pytest.raises(fsb.FscError) as ei:
        ctx.ks_pbs(a, luts, None, b, count=5)                  # range
    assert ei.value.code == 1
    with pytest.raises(fsb.FscError) as ei:
        fsb.Context(fsb.Params.preset("toy").__class__(lwe_dim=48, glwe_dim=2, poly_size=1024, pbs_base_log=23, pbs_level=1,
                                                       ks_base_log=3, ks_level=5, message_modulus=4, carry_modulus=4, acc_bits=64))
    assert ei.value.code == 2
    # the 32-bit accumulator rounds through F2I.S64 (|v| ~ 2^(base_log + 34.4) rms): base_log > 24 is refused there, accepted at 64 bits
    wide = dict(lwe_dim=48, glwe_dim=1, poly_size=2048, pbs_base_log=26, pbs_level=1, ks_base_log=3, ks_level=5, message_modulus=4, carry_modulus=4)
    with pytest.raises(fsb.FscError) as ei:
        fsb.Context(fsb.Params.preset("toy").__class__(acc_bits=32, **wide))
    assert ei.value.code == 2
    fsb.Context(fsb.Params.preset("toy").__class__(acc_bits=64, **wide)).close()
    ctx.close()


@pytest.mark.parametrize("variant", ["stream", "split"])
def test_stream_kernel_noise_statistics(oracle_keys, monkeypatch, variant):
    """the stream kernel, then the split kernel, forced for every width (FSC_PBS_VARIANT): sigma of fresh PBS outputs over
    25 600 real bootstraps of the 2_2 parameter set inside the same budget as the default kernels, zero decode failures."""
    import fhe_sign_b200 as fsb
    monkeypatch.setenv("FSC_PBS_VARIANT", variant)
    K = oracle_keys("2_2_gaussian")
    ctx = fsb.Context(fsb.Params.preset("2_2_gaussian", acc_bits=32))
    ctx.upload_keys(K.bsk, K.ksk)
    assert ctx.pbs_kernel_name() == "pbs_%s_kernel" % variant
    table = (np.arange(16) * 7 + 3) % 16
    luts = ctx.luts_from_tables(table)
    rng = np.random.default_rng(6)
    total, chunk, sq, fails = 25600, 6400, 0.0, 0
    for s in range(total // chunk):
        m = rng.integers(0, 16, chunk).astype(np.uint64)
        out = ctx.apply_lut_host(K.encrypt_msgs(m, seed=12, stream=s * chunk), luts)
        exp = table[m].astype(np.uint64)
        fails += int((K.decrypt_msgs(out) != exp).sum())
        e = _noise(K, out, exp)
        sq += float((e * e).sum())
    sigma = (sq / total) ** 0.5
    print("%s kernel sigma_pbs=2^%.3f over %d bootstraps" % (variant, np.log2(sigma), total))
    assert fails == 0 and sigma < 2.0**-14.0
    ctx.close()


@pytest.mark.parametrize("acc_bits", [64, 32])
def test_noise_over_1e5_bootstraps(gpu_ctx, oracle_keys, acc_bits):
    """north_star: PBS noise within the parameter set's variance bound over >= 1e5 bootstraps.
    Bound: decoding radius 2^-5 at p_fail 2^-64 -> sigma_total <= 2^-5 / 9.2; a fresh PBS output
    (noise level 1) must sit far inside it.  Also: zero decode failures over all 1e5."""
    K, ctx = oracle_keys("2_2_gaussian"), gpu_ctx("2_2_gaussian", acc_bits)
    table = (np.arange(16) * 7 + 3) % 16
    luts = ctx.luts_from_tables(table)
    rng = np.random.default_rng(5)
    total, chunk = 102400, 12800
    sq, fails, worst = 0.0, 0, 0.0
    for s in range(total // chunk):
        m = rng.integers(0, 16, chunk).astype(np.uint64)
        out = ctx.apply_lut_host(K.encrypt_msgs(m, seed=11, stream=s * chunk), luts)
        exp = table[m].astype(np.uint64)
        fails += int((K.decrypt_msgs(out) != exp).sum())
        e = _noise(K, out, exp)
        sq += float((e * e).sum()); worst = max(worst, float(np.abs(e).max()))
    sigma = (sq / total) ** 0.5
    print("acc_bits=%d sigma_pbs=2^%.3f worst=2^%.3f over %d bootstraps" % (acc_bits, np.log2(sigma), np.log2(worst), total))
    assert fails == 0
    assert sigma < 2.0**-14.0          # measured oracle sigma is ~2^-15.1; budget for level-1 noise
    assert worst < 2.0**-6


_ORACLE_SIGMA = {}


@pytest.mark.parametrize("acc_bits", [32, 64])
def test_noise_gate_against_the_oracle_at_full_parameters(gpu_ctx, oracle_keys, acc_bits):
    """SURVEY.md 8d noise gate as written: sigma_pbs^2 of the GPU output <= 1.05 x the CPU oracle's measured
    sigma_pbs^2, at the FULL 2_2 parameter set (n = 834), both on fresh encryptions.  The oracle side runs 8 192
    bootstraps on the host cores (relative standard error of its variance estimate 1.6 %, so 1.05 is three sigma of the
    estimator: the gate is about the kernels, not about sampling luck; seeds are fixed, both sides are deterministic);
    the GPU side runs 32 768.  Covers both accumulator widths: the 32-bit accumulator (default) adds its rounding."""
    from oracle import orc
    K, ctx = oracle_keys("2_2_gaussian"), gpu_ctx("2_2_gaussian", acc_bits)
    table = (np.arange(16) * 7 + 3) % 16
    luts = ctx.luts_from_tables(table)
    rng = np.random.default_rng(21)
    n_cpu, n_gpu = 8192, 32768
    m = rng.integers(0, 16, n_cpu).astype(np.uint64)
    if "var" not in _ORACLE_SIGMA:      # the CPU side is the same for both accumulator widths: run it once
        orc.set_threads(orc.host_cores())
        ref = K.ks_pbs(K.encrypt_msgs(m, seed=13, stream=0), K.make_lut(table), nthreads=orc.host_cores())
        exp = table[m].astype(np.uint64)
        assert (K.decrypt_msgs(ref) == exp).all()
        _ORACLE_SIGMA["var"] = float((_noise(K, ref, exp) ** 2).mean())
    var_cpu = _ORACLE_SIGMA["var"]
    sq = 0.0
    for s in range(n_gpu // 8192):
        m = rng.integers(0, 16, 8192).astype(np.uint64)
        out = ctx.apply_lut_host(K.encrypt_msgs(m, seed=14, stream=s * 8192), luts)
        exp = table[m].astype(np.uint64)
        assert (K.decrypt_msgs(out) == exp).all()
        sq += float((_noise(K, out, exp) ** 2).sum())
    var_gpu = sq / n_gpu
    print("acc_bits=%d  sigma^2 GPU / CPU oracle = %.4f  (GPU 2^%.3f over %d, oracle 2^%.3f over %d)"
          % (acc_bits, var_gpu / var_cpu, 0.5 * np.log2(var_gpu), n_gpu, 0.5 * np.log2(var_cpu), n_cpu))
    assert var_gpu <= 1.05 * var_cpu, (var_gpu, var_cpu)


def test_exact_fourier_key_is_correctly_rounded(oracle_keys, monkeypatch):
    """bsk_exact.cu (direct DFT in double-double arithmetic, the default at key upload) against (a) the same transform in
    80-bit long double on the host: equal to the last bit or two of a double, and (b) the kernels' own f64 FFT conversion
    (FSC_BSK_CONVERT=fft): equal to FFT rounding (1e-16 relative rms) - in both key layouts."""
    import fhe_sign_b200 as fsb
    K = oracle_keys("toy")
    n = K.params.lwe_dim

    def keys(conv):
        if conv:
            monkeypatch.setenv("FSC_BSK_CONVERT", conv)
        else:
            monkeypatch.delenv("FSC_BSK_CONVERT", raising=False)
        ctx = fsb.Context(fsb.Params.preset("toy", acc_bits=32))      # auto: ring layout + stream layout
        ctx.upload_keys(K.bsk, K.ksk)
        out = ctx.debug_fourier_key(False), ctx.debug_fourier_key(True)
        ctx.close()
        return out
    ring_x, stream_x = keys(None)
    ring_f, stream_f = keys("fft")
    for x, f in ((ring_x, ring_f), (stream_x, stream_f)):
        rel = np.sqrt((np.abs(x - f) ** 2).mean() / (np.abs(f) ** 2).mean())
        assert 0 < rel < 1e-15, rel
    # host reference in long double for a few polynomials: X_k = sum_j (a_j + i a_{j+1024}) zeta^(j (4k+1)), lane = k2, slot r: k1 = brev5(r)
    brev5 = lambda v: ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4)
    ld = np.longdouble
    pi = ld("3.14159265358979323846264338327950288")
    e = np.arange(4096).astype(ld)
    cosv, sinv = np.cos(2 * pi * e / 4096), np.sin(2 * pi * e / 4096)
    expo = np.outer(np.arange(1024), 4 * np.arange(1024) + 1) % 4096
    Cm, Sm = cosv[expo], sinv[expo]
    bsk = K.bsk.reshape(n, 4, 2048).view(np.int64)
    worst = 0.0
    for i, g in ((0, 0), (1, 3), (n - 1, 2)):
        a = bsk[i, g]
        zr, zi = a[:1024].astype(ld), a[1024:].astype(ld)
        Xr, Xi = zr @ Cm - zi @ Sm, zr @ Sm + zi @ Cm
        for r in range(32):
            ks = np.arange(32) + 32 * brev5(r)
            want = Xr[ks].astype(np.float64) + 1j * Xi[ks].astype(np.float64)
            got = ring_x[i, r, g, :]
            worst = max(worst, float(np.abs(got - want).max() / np.abs(want).max()))
    assert worst < 4e-16, worst      # a couple of ulps (the host sum itself is only 64-bit accurate); the FFT conversion is ~50x that


@pytest.mark.parametrize("combo", ["7 unit terms", "4x + 3y"])
def test_noise_budget_in_variance_units(gpu_ctx, oracle_keys, combo):
    """Evidence for FSC_RADIX_NOISE=variance (radix.h): the parameter set's budget is a VARIANCE, nu^2 = 25 fresh-PBS
    variances at the input of a lookup, so a linear combination of independent fresh blocks is admissible when
    sum c^2 <= 25 - 7 unit terms (sum |c| = 7 would be refused by the linear rule) or 4x + 3y (sum c^2 = 25 exactly).
    Measured on real ciphertexts at the full 2_2 parameters: 102 400 such combinations of fresh GPU bootstraps go through
    one more lookup with zero decode failures, and the noise that lookup sees (combination + keyswitch, measured under the
    small key on a sample; + the analytic modulus-switch term) leaves z = radius / sigma >= 9.2 (p_fail <= 2^-64: what the
    parameter set is designed for - keyswitch and modulus switch are 99 % of that noise, the combination under 1 %)."""
    import math
    K, ctx = oracle_keys("2_2_gaussian"), gpu_ctx("2_2_gaussian", 32)
    ident = ctx.luts_from_tables(np.arange(16))
    rng = np.random.default_rng(31)
    coefs, tops = ([1] * 7, [2] * 7) if combo.startswith("7") else ([4, 3], [3, 1])      # value <= 14 / 15: inside the 4-bit space
    total, chunk, fails, sq_big = 102400, 12800, 0, 0.0
    small_err = []
    for s in range(total // chunk):
        acc = np.zeros((chunk, 2049), dtype=np.uint64)
        want = np.zeros(chunk, dtype=np.uint64)
        for t, (c, top) in enumerate(zip(coefs, tops)):
            m = rng.integers(0, top + 1, chunk).astype(np.uint64)
            fresh = ctx.apply_lut_host(K.encrypt_msgs(m, seed=40 + t, stream=s * chunk), ident)      # a fresh bootstrap output
            with np.errstate(over="ignore"):
                acc += fresh * np.uint64(c)
            want += m * np.uint64(c)
        out = ctx.apply_lut_host(acc, ident)
        fails += int((K.decrypt_msgs(out) != want).sum())
        e = _noise(K, acc, want)
        sq_big += float((e * e).sum())
        if s == 0:      # noise under the small key after the keyswitch, on 2 048 of them (CPU keyswitch = the GPU's, bit for bit)
            ks = K.keyswitch(acc[:2048])
            small_err = (K.phase_small(ks) - K.encode(want[:2048])).astype(np.int64).astype(np.float64) / 2.0**64
    n = K.params.lwe_dim
    var_in = float((small_err ** 2).mean())
    var_ms = (n / 2 + 1) / (12.0 * 4096.0**2)              # rounding of n/2 + 1 mask terms (binary key) to multiples of 1/4096
    z = 2.0**-6 / math.sqrt(var_in + var_ms)               # decoding radius: half a plaintext step, Delta / 2 = 2^-6 of the torus
    print("%s: sigma combination 2^%.2f (sum c^2 = %d), sigma after keyswitch 2^%.2f, + modulus switch 2^%.2f -> z = %.1f, %d failures in %d"
          % (combo, 0.5 * math.log2(sq_big / total), sum(c * c for c in coefs), 0.5 * math.log2(var_in), 0.5 * math.log2(var_in + var_ms), z, fails, total))
    assert fails == 0
    assert z >= 9.2
