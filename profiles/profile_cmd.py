"""Short, deterministic workload for ncu captures (a few launches of the dominant kernel).
usage: python profiles/profile_cmd.py [ms|bp|ms32|et|bpet] [frames] [frames_per_cta] [threads_per_cta] [codefile]
PAIR=1 in the environment: the two-CTAs-per-SM shape (one lane, 256 threads, 16-bit index tables, 296 CTAs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "ms"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 4 * 8
fpc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
threads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
code = sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt")
ctx = api.Context(code, "", device=0)
ctx.set_tuning(precision=api.F32 if mode == "ms32" else api.F64, frames_per_cta=fpc, threads_per_cta=threads)
if os.environ.get("PAIR") == "1":
    ctx.set_tuning(frames_per_cta=4 if mode == "ms32" else 2, threads_per_cta=256, idx16=2, ctas=296)
dec = "BP" if mode in ("bp", "bpet") else "BP_MS"
snr = float(os.environ.get("SNR", "-4.5"))   # e.g. SNR=3 with mode et: the refill-dominated regime
for i in range(3):
    r = ctx.sim_point("AWGN", snr, seed=0, point=0, frame0=i * frames, nframes=frames, decoding=dec, iterations=50, early_term=(mode in ("et", "bpet")))
    print(r)
print(ctx.stats())
