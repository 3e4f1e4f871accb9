"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
(a) the committed golden vectors produced by the unmodified reference and (b) the C oracle on the
same seeded inputs.  Bars: min-sum — bit-exact posteriors, decisions and iteration counts;
BP — posteriors within 1e-4 relative (north_star), >= 99.99 % identical decisions; integer paths
(BSC/BEC channel, BEC decoder, counters) — bit-exact."""
import numpy as np
import pytest

from conftest import G_FILE, H_FILE

pytestmark = pytest.mark.gpu

BP_RTOL = 1e-4  # tolerance stated by BASELINE.json north_star for BP posterior LLRs


def _cfg(case):
    it, et, seed, n, useg = [int(v) for v in case["cfg"]]
    ch, dec = [str(v) for v in case["names"]]
    return it, bool(et), seed, n, useg, ch, dec, float(case["x"][0])


def _rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-9)


@pytest.mark.parametrize("precision_mode", ["smem", "global"])
def test_minsum_golden_bit_exact(gpu_ctx, golden_cases, precision_mode):
    from libldpc_b200 import api
    gpu_ctx.set_tuning(precision=api.F64, residency=api.SMEM if precision_mode == "smem" else api.GLOBAL, frames_per_cta=0)
    n_checked = 0
    for name, case in golden_cases.items():
        it, et, seed, n, useg, ch, dec, x = _cfg(case)
        if ch == "BEC" or dec != "BP_MS":
            continue
        out, hard, its = gpu_ctx.decode_batch(case["llr_in"], "BP_MS", it, et)
        assert np.array_equal(its, case["iters"]), name
        assert np.array_equal(hard, case["co"]), name
        assert np.array_equal(out.view(np.uint64), case["llr_out"].view(np.uint64)), name  # bit pattern, incl. -0.0
        n_checked += 1
    assert n_checked >= 8
    gpu_ctx.set_tuning(residency=api.AUTO)


def test_bp_golden_within_tolerance(gpu_ctx, golden_cases):
    from libldpc_b200 import api
    gpu_ctx.set_tuning(precision=api.F64, residency=api.AUTO, frames_per_cta=0)
    tot = same = 0
    for name, case in golden_cases.items():
        it, et, seed, n, useg, ch, dec, x = _cfg(case)
        if ch == "BEC" or dec != "BP":
            continue
        out, hard, its = gpu_ctx.decode_batch(case["llr_in"], "BP", it, et)
        assert np.array_equal(its, case["iters"]), name
        tot += hard.size
        same += int((hard == case["co"]).sum())
        assert _rel_err(out, case["llr_out"]).max() < BP_RTOL, name
    assert tot > 0 and same / tot >= 0.9999


def test_bec_golden_bit_exact(gpu_ctx, golden_cases):
    n_checked = 0
    for name, case in golden_cases.items():
        it, et, seed, n, useg, ch, dec, x = _cfg(case)
        if ch != "BEC":
            continue
        out, hard, its = gpu_ctx.decode_bec_batch(case["llr_in"], case["cw"], it, et)
        assert np.array_equal(its, case["iters"]), name
        assert np.array_equal(out, case["llr_out"]), name
        assert np.array_equal(hard, case["co"]), name
        n_checked += 1
    assert n_checked >= 3


@pytest.mark.parametrize("decoding,et,iters", [("BP_MS", True, 50), ("BP_MS", False, 7), ("BP_MS", True, 1), ("BP", True, 50), ("BP", False, 3)])
def test_decode_vs_oracle_seeded(gpu_ctx, oracle_code, decoding, et, iters):
    """Same seeded LLRs (incl. exact zeros, ties, huge values, -0.0) through the oracle and the GPU."""
    rng = np.random.default_rng(1234)
    n = 96
    llr = rng.normal(1.0, 1.6, size=(n, oracle_code.nc))
    llr[:, oracle_code.puncture] = 0.0
    llr[3, :] = np.round(llr[3, :])            # exact ties and zeros
    llr[4, ::7] = -0.0
    llr[5, :] = np.where(rng.random(oracle_code.nc) < 0.2, -1.5, 1.5)  # BSC-like two-valued input
    llr[6, :200] = 99999.9
    llr[7, :] *= 300.0                         # magnitudes of several hundred: the shifted exponentials of the sum-product check node
    llr[8, ::3] = 99999.9                      # checks whose other inputs are all beyond e^-708: its exact (pairwise) path
    llr[9, ::2] *= 1e-9
    ro, rc, ri = oracle_code.decode(llr, iters, et, decoding == "BP_MS")
    out, hard, its = gpu_ctx.decode_batch(llr, decoding, iters, et)
    assert np.array_equal(its, ri)
    if decoding == "BP_MS":
        assert np.array_equal(hard, rc)
        assert np.array_equal(out.view(np.uint64), ro.view(np.uint64))
    else:
        assert (hard == rc).mean() >= 0.9999
        assert _rel_err(out, ro).max() < BP_RTOL


def test_ragged_batches_and_refill(gpu_ctx, oracle_code):
    """Batch sizes around the tile size (1, fpc-1, fpc+1, many) give per-frame identical results."""
    rng = np.random.default_rng(7)
    llr = rng.normal(0.6, 1.3, size=(700, oracle_code.nc))
    llr[:, oracle_code.puncture] = 0.0
    full = gpu_ctx.decode_batch(llr, "BP_MS", 20, True)
    for n in (1, 3, 5, 33, 149 * 4 + 1):
        part = gpu_ctx.decode_batch(llr[:n], "BP_MS", 20, True)
        for a, b in zip(part, full):
            assert np.array_equal(a, b[:n])
    ro, rc, ri = oracle_code.decode(llr[:40], 20, True, True)
    assert np.array_equal(full[2][:40], ri) and np.array_equal(full[1][:40], rc)


def test_large_host_batch_pipeline(gpu_ctx, oracle_code):
    """A batch long enough for the ramped, double-buffered host pipeline (1, 2, 4 ... waves up, steady chunks, down again, ragged
    tail): every frame must come back in its place, equal to the same frame decoded in a small batch and to the oracle."""
    n = 17000 + 37
    cw, llr = gpu_ctx.channel("AWGN", -4.0, seed=11, point=0, frame0=0, n=n)
    out, hard, its = gpu_ctx.decode_batch(llr, "BP_MS", 8, True)
    assert out.shape == (n, oracle_code.nc) and its.shape == (n,)
    rng = np.random.default_rng(5)
    for o in [0, 591, 592, 4143, 4144, n - 700] + list(rng.integers(0, n - 64, 6)):
        part = gpu_ctx.decode_batch(llr[o:o + 64], "BP_MS", 8, True)
        assert np.array_equal(part[0].view(np.uint64), out[o:o + 64].view(np.uint64)), o
        assert np.array_equal(part[1], hard[o:o + 64]) and np.array_equal(part[2], its[o:o + 64]), o
    sel = np.r_[0:8, n - 8:n]
    ro, rc, ri = oracle_code.decode(llr[sel], 8, True, True)
    assert np.array_equal(its[sel], ri) and np.array_equal(hard[sel], rc) and np.array_equal(out[sel].view(np.uint64), ro.view(np.uint64))


def test_channel_kernel_vs_spec(gpu_ctx, oracle_code, oracle_gen):
    """Philox channel: BSC/BEC inputs AND the AWGN LLRs bit-exact with the CPU specification (the normal generator is
    Box-Muller in exactly specified binary32 arithmetic, oracle/ldpc_oracle.c normal_pair_v2).  The context holds a
    generator matrix, so the frames of every channel carry random codewords u*G (the reference's -G)."""
    for ch, x in (("BSC", 0.11), ("BEC", 0.45)):
        cw, llr = gpu_ctx.channel(ch, x, seed=9, point=3, frame0=1 << 33, n=17)
        ocw, ollr = oracle_code.channel_frames(ch, x, 9, 3, 1 << 33, 17, gen=oracle_gen)
        assert np.array_equal(cw, ocw)
        assert np.array_equal(llr, ollr)
        if ch == "BSC":
            assert 0.3 < cw.mean() < 0.7 and not oracle_code.syndrome(cw[0]).any()   # real, non-trivial codewords
    cw, llr = gpu_ctx.channel("AWGN", -4.5, seed=2, point=1, frame0=5, n=64)
    ocw, ollr = oracle_code.channel_frames("AWGN", -4.5, 2, 1, 5, 64, gen=oracle_gen)
    assert np.array_equal(cw, ocw)
    assert np.array_equal(llr.view(np.uint64), ollr.view(np.uint64))
    tx = oracle_code.bit_pos
    z = (llr[:, tx] * 10 ** (4.5 / 10) / 2 - (1 - 2.0 * cw[:, tx])) / np.sqrt(10 ** (4.5 / 10))  # back to standard normal
    gpu_ctx.set_tuning(zero_codeword=1)
    cw0, llr0 = gpu_ctx.channel("AWGN", -4.5, seed=2, point=1, frame0=5, n=8)
    gpu_ctx.set_tuning(zero_codeword=0)
    assert not cw0.any() and np.array_equal(llr0.view(np.uint64), oracle_code.channel_frames("AWGN", -4.5, 2, 1, 5, 8)[1].view(np.uint64))
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    assert np.all(llr[:, oracle_code.puncture] == 0.0)


@pytest.mark.parametrize("ch,x,dec", [("BSC", 0.21, "BP_MS"), ("BSC", 0.16, "BP_MS"), ("BEC", 0.9, "BP"), ("BEC", 0.6, "BP")])
def test_sim_counters_bit_exact_integer_channels(gpu_ctx, oracle_code, oracle_gen, ch, x, dec):
    """Whole pipeline (encode -> channel -> decode -> accounting) on the GPU equals the oracle's frame loop; BSC frames
    carry random codewords through the generator matrix, with and without (zero_codeword) it."""
    n = 600
    g = gpu_ctx.sim_point(ch, x, seed=5, point=2, frame0=100, nframes=n, decoding=dec, iterations=30, early_term=True)
    o = oracle_code.sim_point(ch, x, seed=5, point=2, frame0=100, nframes=n, decoding=dec, iterations=30, early_term=True, threads=8,
                              gen=oracle_gen)
    assert {k: g[k] for k in ("fec", "bec", "frames", "iters")} == o
    if ch == "BSC":
        gpu_ctx.set_tuning(zero_codeword=1)
        g0 = gpu_ctx.sim_point(ch, x, seed=5, point=2, frame0=100, nframes=n, decoding=dec, iterations=30, early_term=True)
        gpu_ctx.set_tuning(zero_codeword=0)
        o0 = oracle_code.sim_point(ch, x, seed=5, point=2, frame0=100, nframes=n, decoding=dec, iterations=30, early_term=True, threads=8)
        assert {k: g0[k] for k in ("fec", "bec", "frames", "iters")} == o0


def test_sim_awgn_minsum_matches_oracle_on_dumped_llrs(gpu_ctx, oracle_code):
    """AWGN: LLRs the GPU channel generates, decoded by the oracle, reproduce the fused kernel's counters."""
    n, x = 400, -4.5
    g = gpu_ctx.sim_point("AWGN", x, seed=1, point=0, frame0=0, nframes=n, decoding="BP_MS", iterations=50, early_term=True)
    cw, llr = gpu_ctx.channel("AWGN", x, seed=1, point=0, frame0=0, n=n)
    assert cw.any()                                                  # random codewords through G
    out, co, its = oracle_code.decode(llr, 50, True, True)
    errs = (co[:, oracle_code.bit_pos] != cw[:, oracle_code.bit_pos]).sum(1)
    assert g["frames"] == n
    assert g["iters"] == int(its.sum())
    assert g["fec"] == int((errs > 0).sum())
    assert g["bec"] == int(errs.sum())


@pytest.mark.parametrize("x,dec,et,iters", [(-4.6, "BP_MS", True, 50), (-5.2, "BP_MS", True, 20), (-4.0, "BP_MS", False, 6), (2.0, "BP_MS", True, 50)])
def test_sim_awgn_counters_equal_oracle_frame_loop(gpu_ctx, oracle_code, oracle_gen, x, dec, et, iters):
    """The AWGN channel is bit-exact with its specification, so the fused GPU pipeline (encode -> channel -> min-sum ->
    accounting) must reproduce the counters of the oracle's own frame loop — no dumped LLRs in between.  Random codewords
    through the generator matrix, and the all-zero word."""
    n = 700
    for zero in (0, 1):
        gpu_ctx.set_tuning(zero_codeword=zero)
        g = gpu_ctx.sim_point("AWGN", x, seed=21, point=4, frame0=1 << 34, nframes=n, decoding=dec, iterations=iters, early_term=et)
        o = oracle_code.sim_point("AWGN", x, seed=21, point=4, frame0=1 << 34, nframes=n, decoding=dec, iterations=iters, early_term=et,
                                  threads=8, gen=None if zero else oracle_gen)
        assert {k: g[k] for k in ("fec", "bec", "frames", "iters")} == o, (x, zero)
    gpu_ctx.set_tuning(zero_codeword=0)


def test_narrow_encodings_int8_float32_and_packed_decisions(gpu_ctx, oracle_code):
    """ldpc_b200_decode_batch_ex: int8 / float32 LLRs in, bit-packed decisions out.  The decoder runs on exactly the doubles
    {int8 * scale} ({(double)float}), so min-sum is bit-identical to the reference decoder (oracle) fed those doubles."""
    rng = np.random.default_rng(5)
    n = 300
    cw, llr = gpu_ctx.channel("AWGN", -4.4, seed=3, point=0, frame0=0, n=n)
    scale = 0.25
    q = np.clip(np.rint(llr / scale), -127, 127).astype(np.int8)
    for src, wide in ((q, q.astype(np.float64) * scale), (llr.astype(np.float32), llr.astype(np.float32).astype(np.float64))):
        for et, iters in ((True, 50), (False, 5)):
            ro, rc, ri = oracle_code.decode(wide, iters, et, True)
            out, hard, bits, its = gpu_ctx.decode_batch_ex(src, "BP_MS", iters, et, scale=scale, want_llr=True, want_hard=True)
            assert np.array_equal(its, ri) and np.array_equal(hard, rc)
            assert np.array_equal(out.view(np.uint64), ro.view(np.uint64))
            assert np.array_equal(gpu_ctx.unpack_bits(bits), rc)
    out, hard, bits, its = gpu_ctx.decode_batch_ex(q[:7], "BP", 50, True, scale=scale)      # bits only, ragged tiny batch
    ro, rc, ri = oracle_code.decode(q[:7].astype(np.float64) * scale, 50, True, False)
    assert out is None and hard is None and np.array_equal(its, ri) and (gpu_ctx.unpack_bits(bits) == rc).mean() >= 0.9999


def test_counters_independent_of_partition(gpu_ctx):
    """Totals depend only on (seed, point, frame range): split launches / CTA counts give the same sums."""
    kw = dict(seed=3, point=1, decoding="BP_MS", iterations=25, early_term=True)
    whole = gpu_ctx.sim_point("AWGN", -4.8, frame0=0, nframes=3000, **kw)
    parts = [gpu_ctx.sim_point("AWGN", -4.8, frame0=f0, nframes=nf, **kw) for f0, nf in ((0, 1000), (1000, 1), (1001, 1999))]
    gpu_ctx.set_tuning(ctas=7)
    few = gpu_ctx.sim_point("AWGN", -4.8, frame0=0, nframes=3000, **kw)
    gpu_ctx.set_tuning(ctas=0)
    for k in ("fec", "bec", "frames", "iters"):
        assert whole[k] == sum(p[k] for p in parts) == few[k]


def test_f32_mode_is_statistically_consistent(gpu_ctx):
    from libldpc_b200 import api
    kw = dict(seed=11, point=0, frame0=0, nframes=20000, decoding="BP_MS", iterations=50, early_term=True)
    a = gpu_ctx.sim_point("AWGN", -4.5, **kw)
    gpu_ctx.set_tuning(precision=api.F32)
    b = gpu_ctx.sim_point("AWGN", -4.5, **kw)
    gpu_ctx.set_tuning(precision=api.F64)
    assert abs(a["fec"] - b["fec"]) <= 0.02 * a["fec"] + 5   # same noise, rounding-level differences only
    assert abs(a["iters"] - b["iters"]) <= 0.01 * a["iters"]


def test_fer_curve_within_binomial_ci(gpu_ctx):
    """FER/BER points agree with the reference CLI's curve (tests/golden/curves.json) within 95 % CIs."""
    import json, os
    from conftest import GOLDEN
    curves = json.load(open(os.path.join(GOLDEN, "curves.json")))
    for name, dec, ch in (("awgn_ms", "BP_MS", "AWGN"), ("awgn_bp", "BP", "AWGN"), ("bsc_ms", "BP_MS", "BSC")):
        for pt in curves[name]["points"]:
            if pt["fer"] < 4e-3 or pt["fec"] < 30:
                continue
            n = int(min(max(400 / pt["fer"], 4000), 120000))
            g = gpu_ctx.sim_point(ch, pt["x"], seed=77, point=0, frame0=0, nframes=n, decoding=dec, iterations=50, early_term=True)
            p_ref, n_ref = pt["fer"], pt["frames"]
            p_gpu = g["fec"] / g["frames"]
            p = (pt["fec"] + g["fec"]) / (n_ref + g["frames"])
            sigma = np.sqrt(p * (1 - p) * (1 / n_ref + 1 / g["frames"]))
            assert abs(p_gpu - p_ref) <= 1.96 * 1.5 * sigma + 1e-12, (name, pt["x"], p_gpu, p_ref, sigma)


def test_reference_abi_and_python_wrapper(built_lib, oracle_code, tmp_path):
    """The six reference symbols via the pyLDPC-compatible wrapper: decode/encode/syndrome/rank/simulate."""
    import time
    from libldpc_b200 import ldpc
    code = ldpc.LDPC(H_FILE, G_FILE)
    assert (code.n, code.m, code.nct, code.mct, code.kct) == (1152, 1024, 1024, 896, 128)
    rng = np.random.default_rng(5)
    llr = rng.normal(1.0, 1.5, size=code.nct)
    full = np.zeros(code.n); full[oracle_code.bit_pos] = llr
    for dec in ("BP_MS", "BP", "BP_MS"):   # no sticky min-sum state between calls
        out, it = code.decode(llr, True, 50, dec)
        ro, rc, ri = oracle_code.decode(full, 50, True, dec == "BP_MS")
        assert it == ri
        if dec == "BP_MS":
            assert np.array_equal(out, ro[oracle_code.bit_pos])
        else:
            assert _rel_err(out, ro[oracle_code.bit_pos]).max() < BP_RTOL
    u = rng.integers(0, 2, code.kct)
    cw = code.encode(u)
    assert cw.shape == (code.nct,)
    assert code.rank() == 1021
    word = np.zeros(code.n, dtype=np.uint8); word[5] = 1
    assert np.array_equal(code.syndrome(word), oracle_code.syndrome(word))
    code.simulate(snr=[-5.5, -4.4, 0.5], iterations=50, decoding="BP_MS", maxFrames=20000, fec=40)
    code.wait(120)
    res = code.get_results()
    assert len(res["fer"]) == 3 and all(f >= 40 for f in res["fec"])
    assert res["fer"][0] > res["fer"][1] > res["fer"][2] > 0


def test_cli_results_file(built_lib, tmp_path):
    import os, subprocess
    from conftest import ROOT
    out = tmp_path / "res.txt"
    cli = os.path.join(ROOT, "libldpc_b200", "ldpcsim")
    r = subprocess.run([cli, H_FILE, str(out), "-6", "-4.9", "0.5", "--decoding", "BP_MS", "--frame-error-count", "30",
                        "--max-frames", "20000", "-t", "4", "-s", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = out.read_text().split("\n")
    assert lines[0] == "snr fer ber frames avg_iter frame_time"
    assert len(lines) == 1 + 3 + 1 and lines[-1] == ""       # header + one line per point + trailing newline
    x, fer, ber, frames, avg_iter, t = lines[1].split()
    assert float(x) == -6.0 and 0.5 < float(fer) <= 1.0 and int(frames) > 0
    assert "N : 1152" in r.stdout and "== Decoder Parameters" in r.stdout and "FEC   |      FRAME" in r.stdout
    bad = subprocess.run([cli, H_FILE, str(out), "3", "1", "0.5"], capture_output=True, text=True)
    assert bad.returncode == 1 and "snr min > snr max" in bad.stdout


@pytest.fixture(scope="module")
def irregular_code():
    """A small irregular code that fits shared memory and exercises every node-update body: check degrees 2..15
    (fixed-degree bodies 2..8 and the generic path), variable degrees 0..81 (bodies 0..16 and the chunked path).
    Committed (tests/golden/irregular_h.txt) together with outputs of the unmodified reference decoder on it."""
    from conftest import H_IRREGULAR
    return H_IRREGULAR


@pytest.mark.parametrize("tmem,idx16", [(0, 0), (1, 0), (0, 2)])
def test_irregular_code_reference_golden(built_lib, irregular_code, irregular_cases, tmem, idx16):
    """CUDA path vs the UNMODIFIED reference decoder's outputs (dump_ref decode) on the irregular code: min-sum bit-pattern
    exact, BP within 1e-4 with identical iteration counts."""
    from libldpc_b200 import api
    ctx = api.Context(irregular_code, "", device=0)
    ctx.set_tuning(precision=api.F64, residency=api.SMEM, tmem=tmem, idx16=idx16)
    for name, case in irregular_cases.items():
        it, et = [int(v) for v in case["cfg"]]
        dec = str(case["names"][0])
        out, hard, its = ctx.decode_batch(case["llr_in"], dec, it, bool(et))
        assert np.array_equal(its, case["iters"]), name
        if dec == "BP_MS":
            assert np.array_equal(hard, case["co"]), name
            assert np.array_equal(out.view(np.uint64), case["llr_out"].view(np.uint64)), name
        else:
            assert (hard == case["co"]).mean() >= 0.9999, name
            assert _rel_err(out, case["llr_out"]).max() < BP_RTOL, name
    ctx.set_tuning(residency=api.GLOBAL)
    for name in ("ms_et50_easy", "ms_bsc", "ms_noet9"):
        case = irregular_cases[name]
        it, et = [int(v) for v in case["cfg"]]
        out, hard, its = ctx.decode_batch(case["llr_in"], "BP_MS", it, bool(et))
        assert np.array_equal(its, case["iters"]) and np.array_equal(out.view(np.uint64), case["llr_out"].view(np.uint64)), name
    ctx.close()


@pytest.mark.parametrize("tmem,idx16", [(0, 0), (1, 0), (0, 2)])
@pytest.mark.parametrize("decoding,et,iters", [("BP_MS", True, 40), ("BP_MS", False, 9), ("BP", True, 25), ("BP", False, 3)])
def test_irregular_code_all_bodies(built_lib, irregular_code, decoding, et, iters, tmem, idx16):
    from libldpc_b200 import api
    from oracle import oracle as O
    ctx = api.Context(irregular_code, "", device=0)
    oc = O.Code(irregular_code)
    ctx.set_tuning(precision=api.F64, residency=api.SMEM, tmem=tmem, idx16=idx16)
    rng = np.random.default_rng(99)
    llr = rng.normal(1.4, 1.9, size=(37, oc.nc))
    llr[:, oc.puncture] = 0.0
    llr[:, oc.shorten] = 99999.9
    llr[2, ::5] = -0.0
    ro, rc, ri = oc.decode(llr, iters, et, decoding == "BP_MS")
    if O.ref_available():   # the comparand is the reference itself wherever its build travelled along (it does to the GPU box)
        r = O.ref_decode(irregular_code, decoding, iters, et, llr, tmp="/tmp/orc_gpu_irr")
        assert np.array_equal(ro.view(np.uint64), r["llr_out"].view(np.uint64)) and np.array_equal(rc, r["co"]) and np.array_equal(ri, r["iters"])
        ro, rc, ri = r["llr_out"], r["co"], r["iters"]
    out, hard, its = ctx.decode_batch(llr, decoding, iters, et)
    assert np.array_equal(its, ri)
    if decoding == "BP_MS":
        assert np.array_equal(hard, rc)
        assert np.array_equal(out.view(np.uint64), ro.view(np.uint64))
    else:
        assert (hard == rc).mean() >= 0.9999
        assert _rel_err(out, ro).max() < BP_RTOL
    g = ctx.sim_point("BSC", 0.04, seed=3, point=0, frame0=0, nframes=300, decoding=decoding, iterations=iters, early_term=et)
    o = oc.sim_point("BSC", 0.04, seed=3, point=0, frame0=0, nframes=300, decoding=decoding, iterations=iters, early_term=et, threads=8)
    if decoding == "BP_MS":
        assert {k: g[k] for k in ("fec", "bec", "frames", "iters")} == o
    ctx.close()


def test_cli_with_generator_matrix(built_lib, tmp_path):
    """-G: the sweep transmits random codewords; the curve stays that of the code (decoders are symmetric)."""
    import os, subprocess
    from conftest import ROOT
    cli = os.path.join(ROOT, "libldpc_b200", "ldpcsim")
    outs = []
    for extra in ([], ["-G", G_FILE]):
        out = tmp_path / ("res%d.txt" % len(outs))
        r = subprocess.run([cli, H_FILE, str(out), "-5.5", "-5.4", "0.5", "--decoding", "BP_MS", "--frame-error-count", "400",
                            "--max-frames", "40000", "-s", "3"] + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        x, fer, ber, frames, avg_iter, t = out.read_text().split("\n")[1].split()
        outs.append((float(fer), int(frames)))
    (f0, n0), (f1, n1) = outs
    p = (f0 * n0 + f1 * n1) / (n0 + n1)
    assert abs(f0 - f1) < 4 * np.sqrt(p * (1 - p) * (1 / n0 + 1 / n1)) + 1e-9, outs


def test_full_size_counters_identical_across_kernel_variants(gpu_ctx):
    """At the bench's step size (303,104 frames, 50 iterations): the shared-memory kernel with the TMEM mirror, without it,
    and the global-residency kernel do the same IEEE operations in the same order, so the four counters must be IDENTICAL;
    early termination on/off must agree on everything but the iteration count for frames that converge."""
    from libldpc_b200 import api
    n = 148 * 4 * 512
    kw = dict(seed=21, point=4, frame0=1 << 40, nframes=n, decoding="BP_MS", iterations=50)
    res = {}
    shapes = {}
    auto = dict(frames_per_cta=0, threads_per_cta=0, ctas=0)
    for name, t in (("smem+tmem", dict(residency=api.SMEM, tmem=0, idx16=0, **auto)),     # automatic (whichever shape won the timed trial)
                    ("smem+tmem32", dict(residency=api.SMEM, tmem=0, idx16=1, **auto)),   # one CTA per SM, 32-bit index tables
                    ("pair16", dict(residency=api.SMEM, tmem=0, idx16=2, frames_per_cta=2, threads_per_cta=256, ctas=296)),  # two per SM, 16-bit tables
                    ("smem", dict(residency=api.SMEM, tmem=1, idx16=0, **auto)), ("global", dict(residency=api.GLOBAL, tmem=0, idx16=0, **auto))):
        gpu_ctx.set_tuning(precision=api.F64, **t)
        res[name] = {et: gpu_ctx.sim_point("AWGN", -4.2, early_term=et, **kw) for et in (True, False)}
        st = gpu_ctx.stats()
        shapes[name] = (st["frames_per_cta"], st["threads_per_cta"], st["ctas"])
    gpu_ctx.set_tuning(residency=api.AUTO, tmem=0, idx16=0, **auto)
    assert shapes["pair16"] == (2, 256, 296) and shapes["smem+tmem32"] == (4, 512, 148), shapes
    for et in (True, False):
        ref = {k: res["smem+tmem"][et][k] for k in ("fec", "bec", "frames", "iters")}
        assert ref["frames"] == n
        for name in ("smem+tmem32", "pair16", "smem", "global"):
            assert {k: res[name][et][k] for k in ("fec", "bec", "frames", "iters")} == ref, (name, et)
    a, b = res["smem+tmem"][True], res["smem+tmem"][False]
    assert b["iters"] == 50 * n and a["iters"] < b["iters"]
    # a frame that never satisfies all checks is decoded identically; one that does usually, but not always, keeps its
    # decisions when min-sum keeps iterating past the codeword
    assert abs(a["fec"] - b["fec"]) <= 0.03 * b["fec"] + 5


def test_bp_counters_identical_across_kernel_shapes(gpu_ctx):
    """fp64 sum-product: one CTA x 384 threads, two CTAs x 192 threads with 16-bit tables (the bench shape), without the TMEM mirror,
    and global residency (per-lane evaluation, bodies inlined) perform the same operations per frame: identical counters."""
    from libldpc_b200 import api
    n = 148 * 4 * 16
    auto = dict(frames_per_cta=0, threads_per_cta=0, ctas=0)
    res, shapes = {}, {}
    for name, t in (("one", dict(residency=api.SMEM, tmem=0, idx16=1, **auto)),
                    ("pair", dict(residency=api.SMEM, tmem=0, idx16=2, frames_per_cta=2, threads_per_cta=192, ctas=296)),
                    ("no-mirror", dict(residency=api.SMEM, tmem=1, idx16=0, **auto)), ("global", dict(residency=api.GLOBAL, tmem=0, idx16=0, **auto))):
        gpu_ctx.set_tuning(precision=api.F64, **t)
        res[name] = {et: gpu_ctx.sim_point("AWGN", -4.6, seed=5, point=2, frame0=12345, nframes=n, decoding="BP", iterations=25, early_term=et) for et in (True, False)}
        st = gpu_ctx.stats()
        shapes[name] = (st["frames_per_cta"], st["threads_per_cta"], st["ctas"])
    gpu_ctx.set_tuning(residency=api.AUTO, tmem=0, idx16=0, **auto)
    assert shapes["one"] == (4, 384, 148) and shapes["pair"] == (2, 192, 296), shapes
    for et in (True, False):
        ref = {k: res["one"][et][k] for k in ("fec", "bec", "frames", "iters")}
        assert ref["frames"] == n and ref["fec"] > 0
        for name in ("pair", "no-mirror", "global"):
            assert {k: res[name][et][k] for k in ("fec", "bec", "frames", "iters")} == ref, (name, et)


@pytest.mark.parametrize("et,iters,compat", [(True, 50, 1), (False, 7, 1), (True, 20, 0), (True, 1, 1)])
def test_bec_bit_sliced_sweep_equals_bytewise_and_oracle(gpu_ctx, oracle_code, et, iters, compat):
    """Erasure sweep: the bit-sliced kernel (32 frames per word, shared memory) against the byte-wise kernel (forced by
    residency=GLOBAL) on ragged frame counts, and both against the oracle's frame loop."""
    from libldpc_b200 import api
    gpu_ctx.set_tuning(bec_deg1_compat=compat, zero_codeword=1)     # the lean all-zero-codeword instantiation
    for eps, n in ((0.88, 1000), (0.6, 333), (0.93, 65)):
        kw = dict(seed=8, point=1, frame0=77, nframes=n, decoding="BP", iterations=iters, early_term=et)
        gpu_ctx.set_tuning(residency=api.AUTO)
        a = gpu_ctx.sim_point("BEC", eps, **kw)
        st = gpu_ctx.stats()
        assert st["residency"] == api.SMEM and st["frames_per_cta"] % 32 == 0      # the bit-sliced kernel ran
        gpu_ctx.set_tuning(residency=api.GLOBAL)
        b = gpu_ctx.sim_point("BEC", eps, **kw)
        gpu_ctx.set_tuning(residency=api.AUTO)
        o = oracle_code.sim_point("BEC", eps, bec_deg1_compat=bool(compat), threads=8, **kw)
        assert {k: a[k] for k in ("fec", "bec", "frames", "iters")} == {k: b[k] for k in ("fec", "bec", "frames", "iters")} == o, (eps, n)
    gpu_ctx.set_tuning(bec_deg1_compat=1, zero_codeword=0)


def test_error_log_names_the_failing_frames(gpu_ctx, oracle_code):
    """Per-error diagnostics log: the logged (global frame, bit errors, iterations) are exactly the frames the oracle finds
    in error on the LLRs the GPU channel dumps, and a logged frame can be replayed from its index alone."""
    n, x, f0 = 800, -4.6, 5000
    kw = dict(seed=13, point=2, frame0=f0, nframes=n, decoding="BP_MS", iterations=40, early_term=True)
    cnt, log, n_err = gpu_ctx.sim_point_log("AWGN", x, capacity=n, **kw)
    plain = gpu_ctx.sim_point("AWGN", x, **kw)
    assert cnt == {k: plain[k] for k in ("fec", "bec", "frames", "iters")} and n_err == cnt["fec"] == len(log) > 10
    cw, llr = gpu_ctx.channel("AWGN", x, seed=13, point=2, frame0=f0, n=n)
    out, co, its = oracle_code.decode(llr, 40, True, True)
    errs = (co[:, oracle_code.bit_pos] != cw[:, oracle_code.bit_pos]).sum(1)
    want = sorted((f0 + int(i), int(errs[i]), int(its[i])) for i in np.nonzero(errs)[0])
    assert log == want
    rep = gpu_ctx.error_report("AWGN", x, 13, 2, log[0][0], decoding="BP_MS", iterations=40, early_term=True)
    assert rep["hamming_distance"] == log[0][1] and rep["iterations"] == log[0][2]
    i = log[0][0] - f0
    assert rep["failed_bits"] == [int(b) for b in oracle_code.bit_pos[co[i][oracle_code.bit_pos] != cw[i][oracle_code.bit_pos]]]
    assert rep["syndrome_weight"] == int(oracle_code.syndrome(co[i]).sum()) == len(rep["failed_checks"])
    cnt2, log2, n2 = gpu_ctx.sim_point_log("AWGN", x, capacity=5, **kw)        # a short buffer truncates, the count does not
    assert n2 == n_err and len(log2) == 5 and set(log2) <= set(log)


@pytest.mark.parametrize("compat", [1, 0])
@pytest.mark.parametrize("zero", [0, 1])
def test_bec_sweep_with_generator_and_error_log(gpu_ctx, oracle_code, oracle_gen, zero, compat):
    """Erasure channel with -G (random codewords u*G, src/sim/channel.cpp:177-191) and the per-error log: the bit-sliced sweep
    kernel (known / wrong bit-planes), the byte-wise kernel and the oracle's frame loop give the same counters; the log names
    exactly the frames the oracle finds in error, with their bit-error and iteration counts; a logged frame replays."""
    from libldpc_b200 import api
    gpu_ctx.set_tuning(zero_codeword=zero, bec_deg1_compat=compat)
    gen = None if zero else oracle_gen
    for eps, n, et, iters in ((0.93, 700, True, 50), (0.88, 2100, True, 30), (0.9, 333, False, 6)):
        kw = dict(seed=17, point=3, frame0=1 << 32, nframes=n, decoding="BP", iterations=iters, early_term=et)
        o = oracle_code.sim_point("BEC", eps, bec_deg1_compat=bool(compat), threads=8, gen=gen, **kw)
        a = gpu_ctx.sim_point("BEC", eps, **kw)
        gpu_ctx.set_tuning(residency=api.GLOBAL)
        b = gpu_ctx.sim_point("BEC", eps, **kw)
        cntb, logb, nb = gpu_ctx.sim_point_log("BEC", eps, capacity=n, **kw)
        gpu_ctx.set_tuning(residency=api.AUTO)
        assert {k: a[k] for k in o} == {k: b[k] for k in o} == o, (eps, zero, compat)
        cnt, log, n_err = gpu_ctx.sim_point_log("BEC", eps, capacity=n, **kw)
        assert cnt == o == cntb and n_err == o["fec"] == len(log) == nb and log == logb
        cw, u8 = gpu_ctx.channel("BEC", eps, seed=17, point=3, frame0=1 << 32, n=n)
        ocw, ou8 = oracle_code.channel_frames("BEC", eps, 17, 3, 1 << 32, n, gen=gen)
        assert np.array_equal(cw, ocw) and np.array_equal(u8, ou8) and (zero or 0.3 < cw.mean() < 0.7)
        out, co, its = oracle_code.decode_bec(u8, cw, iters, et, bool(compat))
        errs = (co[:, oracle_code.bit_pos] != cw[:, oracle_code.bit_pos]).sum(1)
        want = sorted(((1 << 32) + int(i), int(errs[i]), int(its[i])) for i in np.nonzero(errs)[0])
        assert log == want
        if log:
            rep = gpu_ctx.error_report("BEC", eps, 17, 3, log[0][0], iterations=iters, early_term=et)
            assert rep["hamming_distance"] == log[0][1] and rep["iterations"] == log[0][2]
    gpu_ctx.set_tuning(zero_codeword=0, bec_deg1_compat=1)


def test_cli_multi_gpu_gives_identical_results(built_lib, tmp_path):
    """ldpcsim --gpus N shards every round over N GPUs of one process: same counters, hence the same results file
    (all columns but the frame time), as one GPU."""
    import os, subprocess
    from conftest import ROOT
    if built_lib.ldpc_b200_device_count() < 2:
        pytest.skip("needs two GPUs")
    cli = os.path.join(ROOT, "libldpc_b200", "ldpcsim")
    rows = []
    for g in (1, 2):
        out = tmp_path / f"res_g{g}.txt"
        r = subprocess.run([cli, H_FILE, str(out), "-5.5", "-4.4", "0.5", "--decoding", "BP_MS", "--frame-error-count", "300",
                            "--max-frames", "200000", "-s", "4", "--gpus", str(g)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        rows.append([l.split()[:5] for l in out.read_text().split("\n")[1:] if l.strip()])
    assert rows[0] == rows[1] and len(rows[0]) == 3


def test_cli_shards_on_one_gpu_give_identical_results(built_lib, tmp_path):
    """R-independence on a one-GPU lease: the CUDA path run as two (and three) shard 'ranks' on the SAME device
    (--devices 0,0: contiguous split of every round, one context + host thread per shard) produces the counters, hence the
    results file (all columns but the frame time), of the single-range run — sum-product and min-sum, AWGN and BEC."""
    import os, subprocess
    from conftest import ROOT
    cli = os.path.join(ROOT, "libldpc_b200", "ldpcsim")
    for extra in (["-5.5", "-4.4", "0.5", "--decoding", "BP_MS", "--frame-error-count", "300", "--max-frames", "200000"],
                  ["-5.5", "-4.9", "0.5", "--decoding", "BP", "--frame-error-count", "100", "--max-frames", "60000", "-i", "20"],
                  ["0.80", "0.9", "0.05", "--channel", "BEC", "--frame-error-count", "200", "--max-frames", "200000"]):
        rows = []
        for devs in ("0", "0,0", "0,0,0"):
            out = tmp_path / f"res_{len(rows)}.txt"
            r = subprocess.run([cli, H_FILE, str(out)] + extra + ["-s", "4", "--devices", devs], capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout + r.stderr
            rows.append([l.split()[:5] for l in out.read_text().split("\n")[1:] if l.strip()])
        assert rows[0] == rows[1] == rows[2] and len(rows[0]) >= 1, rows


@pytest.mark.parametrize("prec", ["f64"])
@pytest.mark.parametrize("decoding,et,iters", [("BP_MS", True, 50), ("BP_MS", False, 7), ("BP", True, 30), ("BP", False, 3)])
def test_layered_schedule_equals_its_specification(gpu_ctx, oracle_code, decoding, et, iters, prec):
    """Opt-in layered schedule (legacy tree gpu/device/kernel.cpp:52-75 with the live decoder's check rules) against its
    specification oracle.decode_layered: min-sum bit-pattern exact, sum-product within 1e-4, identical iteration counts; built-in
    and caller-supplied layers; fused sweep counters = specification on the dumped LLRs; about half the flooding iterations."""
    from libldpc_b200 import api
    gpu_ctx.set_tuning(schedule=api.LAYERED, zero_codeword=1)
    try:
        for layers in (None, [list(map(int, l[::-1])) for l in oracle_code.auto_layers()[::-1]]):
            gpu_ctx.set_layers(layers)
            use = gpu_ctx.layers()
            cw, llr = gpu_ctx.channel("AWGN", -4.4, seed=31, point=0, frame0=0, n=101)
            llr[3, ::9] = -0.0
            ro, rc, ri = oracle_code.decode_layered(llr, use, iters, et, decoding == "BP_MS")
            for fpc in (0, 4):
                gpu_ctx.set_tuning(frames_per_cta=fpc)
                out, hard, its = gpu_ctx.decode_batch(llr, decoding, iters, et)
                assert np.array_equal(its, ri)
                if decoding == "BP_MS":
                    assert np.array_equal(hard, rc) and np.array_equal(out.view(np.uint64), ro.view(np.uint64))
                else:
                    assert (hard == rc).mean() >= 0.9999 and _rel_err(out, ro).max() < BP_RTOL
            gpu_ctx.set_tuning(frames_per_cta=0)
        gpu_ctx.set_layers(None)
        n = 500
        g = gpu_ctx.sim_point("AWGN", -4.6, seed=5, point=1, frame0=77, nframes=n, decoding=decoding, iterations=iters, early_term=et)
        cw, llr = gpu_ctx.channel("AWGN", -4.6, seed=5, point=1, frame0=77, n=n)
        ro, rc, ri = oracle_code.decode_layered(llr, gpu_ctx.layers(), iters, et, decoding == "BP_MS")
        errs = (rc[:, oracle_code.bit_pos] != 0).sum(1)
        if decoding == "BP_MS":
            assert g["frames"] == n and g["iters"] == int(ri.sum()) and g["fec"] == int((errs > 0).sum()) and g["bec"] == int(errs.sum())
        if et and iters == 50 and decoding == "BP_MS":
            fo, fc, fi = oracle_code.decode(llr, iters, et, True)
            conv = (ri < iters) & (fi < iters)
            assert ri[conv].mean() < 0.7 * fi[conv].mean()          # the point of the schedule
    finally:
        gpu_ctx.set_layers(None)
        gpu_ctx.set_tuning(schedule=api.FLOODING, zero_codeword=0, frames_per_cta=0)


def test_layered_schedule_on_the_bg1_shaped_code(built_lib):
    """Layered min-sum on the quasi-cyclic BG1-shaped code (the case the schedule exists for): bit-exact against the specification."""
    from conftest import large_code_files
    from libldpc_b200 import api
    from oracle import oracle as O
    path = large_code_files()["bg1"]
    ctx = api.Context(path, "", device=0)
    oc = O.Code(path)
    ctx.set_tuning(schedule=api.LAYERED)
    layers = ctx.layers()
    assert oc.layers_valid(layers) and len(layers) <= 46
    cw, llr = oc.channel_frames("AWGN", -0.6, 9, 0, 0, 9)
    for et, iters, q in ((True, 50, 0), (False, 4, 0), (True, 50, 48)):
        ctx.set_tuning(layered_ms_scale64=q)
        ro, rc, ri = oc.decode_layered(llr, layers, iters, et, True, ms_scale=(q or 64) / 64.0)
        out, hard, its = ctx.decode_batch(llr, "BP_MS", iters, et)
        assert np.array_equal(its, ri) and np.array_equal(hard, rc) and np.array_equal(out.view(np.uint64), ro.view(np.uint64))
    assert (ri < 50).sum() >= 4      # normalised layered min-sum converges where the plain one does not (all nine frames stay at 50)
    ctx.close()


@pytest.mark.parametrize("M,snr", [(4, 1.5), (16, 10.0)])
def test_mask_bit_metric_channel_and_sweep(gpu_ctx, oracle_code, M, snr):
    """Higher-order modulation (legacy tree: M-ASK + bit-metric decoding, gpu/device/kernel.cpp:141-219): the GPU channel against
    its specification (scrambling bits identical, LLRs to 1e-9 — library exp / log on both sides), default and caller-supplied
    labels / bit mapper, and the fused sweep's counters against the oracle decoding the dumped LLRs."""
    rng = np.random.default_rng(M)
    lab, bm = oracle_code.default_modulation(M)
    try:
        for labels, mapper in ((None, None), (lab[::-1].copy(), rng.permutation(bm.reshape(-1)).reshape(bm.shape))):
            gpu_ctx.set_modulation(M, labels, mapper)
            l_use, m_use = (lab, bm) if labels is None else (labels, mapper)
            cw, llr = gpu_ctx.channel("AWGN", snr, seed=12, point=2, frame0=1 << 35, n=40)
            ocw, ollr = oracle_code.channel_frames_ask(M, l_use, m_use, snr, 12, 2, 1 << 35, 40)
            assert np.array_equal(cw, ocw)
            assert np.allclose(llr, ollr, rtol=1e-9, atol=1e-9)
            assert np.all(llr[:, oracle_code.puncture] == 0.0) and 0.4 < cw[:, oracle_code.bit_pos].mean() < 0.6
        gpu_ctx.set_modulation(M)
        n = 400
        g = gpu_ctx.sim_point("AWGN", snr - 1.0, seed=3, point=0, frame0=9, nframes=n, decoding="BP_MS", iterations=30, early_term=True)
        cw, llr = gpu_ctx.channel("AWGN", snr - 1.0, seed=3, point=0, frame0=9, n=n)
        ro, rc, ri = oracle_code.decode(llr, 30, True, True)
        errs = (rc[:, oracle_code.bit_pos] != 0).sum(1)
        assert g["frames"] == n and g["iters"] == int(ri.sum()) and g["fec"] == int((errs > 0).sum()) and g["bec"] == int(errs.sum())
        assert 0 < g["iters"] < 30 * n
    finally:
        gpu_ctx.set_modulation(2)
    gpu_ctx.set_tuning(zero_codeword=1)                                               # back to the reference's BPSK channel
    cw, llr = gpu_ctx.channel("AWGN", -4.5, seed=2, point=1, frame0=5, n=4)
    gpu_ctx.set_tuning(zero_codeword=0)
    assert np.array_equal(llr.view(np.uint64), oracle_code.channel_frames("AWGN", -4.5, 2, 1, 5, 4)[1].view(np.uint64))
