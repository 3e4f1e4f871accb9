"""FER / BER vs SNR side by side: the UNMODIFIED reference CLI (oracle/_ref/ldpcsim_ref, all host cores) and this library's sweep
driver on the same points -> markdown.  usage: python profiles/fer_match.py"""
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api  # noqa: E402

H = os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt")
REF = os.path.join(ROOT, "oracle", "_ref", "ldpcsim_ref")
threads = os.cpu_count() or 1
ctx = api.Context(H, "", device=0)


def ref_point(channel, decoding, x, fec, max_frames):
    out = f"/tmp/fer_ref_{os.getpid()}.txt"
    if os.path.exists(out):
        os.remove(out)
    cmd = [REF, H, out, str(x), str(x + 1e-6), "1", "--channel", channel, "--decoding", decoding, "-i", "50", "--frame-error-count", str(fec),
           "--max-frames", str(max_frames), "-t", str(threads), "-s", "1"]
    t0 = time.perf_counter()
    subprocess.run(cmd, stdout=subprocess.DEVNULL, check=True)
    dt = time.perf_counter() - t0
    rows = [l.split() for l in open(out).read().splitlines()[1:] if l.strip()] if os.path.exists(out) else []
    if not rows:
        return None, dt
    _, fer, ber, frames, avg_it, _ = rows[-1]
    return (float(fer), float(ber), int(frames), float(avg_it)), dt


def ci(p, n):
    return 1.96 * math.sqrt(max(p * (1 - p), 1e-30) / n)


print(f"reference: oracle/_ref/ldpcsim_ref -t {threads}; this library: sweep driver on one B200 (fp64).  +/- = 95 % binomial half-width.\n")
plan = [("AWGN", "BP_MS", [-5.5, -5.0, -4.5, -4.0, -3.5], 3000, 20000), ("AWGN", "BP", [-6.0, -5.5, -5.0, -4.5, -4.0], 2000, 20000),
        ("BSC", "BP_MS", [0.22, 0.21, 0.20, 0.19], 3000, 20000)]
for channel, decoding, xs, fec_ref, fec_gpu in plan:
    print(f"### {channel}, {decoding}, -i 50, early termination\n")
    print("| x | ref FER | ref frames | ref avg it | ref s | B200 FER | B200 frames | B200 avg it | B200 s | FER agree | BER ref / B200 |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for x in xs:
        ref, dt_ref = ref_point(channel, decoding, x, fec_ref, 30_000_000)
        t0 = time.perf_counter()
        r = ctx.simulate([x, x + 1e-9, 1.0], channel=channel, decoding=decoding, iterations=50, early_term=True, seed=3, max_frames=300_000_000, fec=fec_gpu)
        dt = time.perf_counter() - t0
        g = (float(r["fer"][0]), float(r["ber"][0]), int(r["frames"][0]), float(r["avg_iter"][0]))
        if ref is None:
            print(f"| {x:g} | no errors | - | - | {dt_ref:.1f} | {g[0]:.3e} | {g[2]} | {g[3]:.2f} | {dt:.2f} | - | - |", flush=True)
            continue
        ok = abs(ref[0] - g[0]) <= ci(ref[0], ref[2]) + ci(g[0], g[2])
        print(f"| {x:g} | {ref[0]:.3e} +/- {ci(ref[0], ref[2]):.1e} | {ref[2]} | {ref[3]:.2f} | {dt_ref:.1f} | {g[0]:.3e} +/- {ci(g[0], g[2]):.1e} | {g[2]} | "
              f"{g[3]:.2f} | {dt:.2f} | {'yes' if ok else 'NO'} | {ref[1]:.3e} / {g[1]:.3e} |", flush=True)
    print()
