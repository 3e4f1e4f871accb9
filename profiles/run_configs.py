"""The five BASELINE.json configurations through the sweep driver (the code path of `ldpcsim` / simulate()) on one B200:
frames, FER, average iterations and throughput per sweep point -> markdown on stdout.
usage: python profiles/run_configs.py [c1 c2 c3 c4 c5 ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "codes"))
import gen_codes  # noqa: E402
from libldpc_b200 import api  # noqa: E402

H = os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt")
big = gen_codes.ensure()
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]


def sweep(title, code, xs, **kw):
    ctx = api.Context(code, "", device=0)
    print(f"\n### {title}\n")
    print("| x | frames (at the last frame error) | frame errors | FER | BER | avg iters | wall s | frames/s (all frames decoded) | coded Gb/s | G edge-it/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for x in xs:
        t0 = time.perf_counter()
        ctx.stats(reset=True)
        r = ctx.simulate([x, x + 1e-9, 1.0], **kw)
        dt = time.perf_counter() - t0
        st = ctx.stats()
        if len(r["frames"]) == 0:
            frames, fec, fer, ber, avg = st["frames"], 0, 0.0, 0.0, float("nan")
        else:
            frames, fec, fer, ber, avg = int(r["frames"][0]), int(r["fec"][0]), float(r["fer"][0]), float(r["ber"][0]), float(r["avg_iter"][0])
        run = st["frames"]   # frames actually decoded; like the reference, the result arrays hold the counts at the LAST frame error
        print(f"| {x:g} | {frames} | {fec} | {fer:.3e} | {ber:.3e} | {avg:.2f} | {dt:.2f} | {run / dt:.3e} | {run * ctx.nct / dt / 1e9:.3f} | "
              f"{st['edge_iterations'] / dt / 1e9:.1f} |", flush=True)
    print(f"\nkernel configuration of the last point: {st['frames_per_cta']} frames/CTA, {st['threads_per_cta']} threads, {st['ctas']} CTAs, "
          f"residency {'shared memory' if st['residency'] == 1 else 'global'}")
    ctx.close()


if "c1" in which:
    sweep("C1: h.txt, AWGN, BP (sum-product, fp64), -i 50, early termination, max-frames 2e6 per point", H,
          [0, 0.5, 1, 1.5, 2, 2.5, 3, 3.5], channel="AWGN", decoding="BP", iterations=50, early_term=True, max_frames=2_000_000, fec=50)
    sweep("C1 waterfall: same, -7 ... -4 dB, --frame-error-count 200", H,
          [-7, -6.5, -6, -5.5, -5, -4.5, -4], channel="AWGN", decoding="BP", iterations=50, early_term=True, max_frames=3_000_000, fec=200)
if "c2" in which:
    sweep("C2: h.txt, AWGN, BP_MS (min-sum, fp64), -i 50, early termination, --frame-error-count 200", H,
          [-6, -5.5, -5, -4.5, -4, -3.5], channel="AWGN", decoding="BP_MS", iterations=50, early_term=True, max_frames=30_000_000, fec=200)
    sweep("C2 high SNR: same, 0 ... 3.5 dB, max-frames 2e7 per point", H,
          [0, 1, 2, 3, 3.5], channel="AWGN", decoding="BP_MS", iterations=50, early_term=True, max_frames=20_000_000, fec=50)
if "c3" in which:
    for et in (True, False):
        sweep(f"C3: BG1-shaped code Z=384 (26112 x 17664), AWGN, BP_MS, -i 50, early termination {'on' if et else 'off'}", big["bg1"],
              [-1.0, -0.75, -0.5, -0.25] if et else [-0.5], channel="AWGN", decoding="BP_MS", iterations=50, early_term=et, max_frames=60_000, fec=100)
if "c4" in which:
    sweep("C4: DVB-S2-shaped code n=64800 r=1/2, AWGN, BP (sum-product, fp64), -i 50, --no-early-term", big["dvbs2"],
          [1.0], channel="AWGN", decoding="BP", iterations=50, early_term=False, max_frames=24_000, fec=10 ** 9)
if "c5" in which:
    sweep("C5: h.txt, BSC, BP_MS, -i 50, --frame-error-count 100, max-frames 3e8", H,
          [0.20, 0.18, 0.16, 0.14, 0.12, 0.10, 0.08], channel="BSC", decoding="BP_MS", iterations=50, early_term=True, max_frames=300_000_000, fec=100)
    sweep("C5: h.txt, BEC, erasure decoder, -i 50, --frame-error-count 100, max-frames 1e8", H,
          [0.90, 0.85, 0.80, 0.75, 0.70], channel="BEC", decoding="BP", iterations=50, early_term=True, max_frames=100_000_000, fec=100)
