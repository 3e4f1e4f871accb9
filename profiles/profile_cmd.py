"""Short, deterministic workload for ncu captures (a few launches of the dominant kernel).
usage: python profiles/profile_cmd.py [ms|bp|ms32|et] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "ms"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 4 * 8
ctx = api.Context(os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt"), "", device=0)
if mode == "ms32":
    ctx.set_tuning(precision=api.F32)
dec = "BP" if mode == "bp" else "BP_MS"
for i in range(4):
    r = ctx.sim_point("AWGN", -4.5, seed=0, point=0, frame0=i * frames, nframes=frames, decoding=dec, iterations=50, early_term=(mode == "et"))
    print(r)
print(ctx.stats())
