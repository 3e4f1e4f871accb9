"""Short global-residency workload for ncu captures.  usage: python profiles/profile_large.py codefile frames snr [lanes] [threads] [BP|BP_MS]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api
code, frames, snr = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
threads = int(sys.argv[5]) if len(sys.argv) > 5 else 512
dec = sys.argv[6] if len(sys.argv) > 6 else "BP_MS"
ctx = api.Context(code, "", device=0)
ctx.set_tuning(precision=api.F64, residency=api.GLOBAL, frames_per_cta=2 * lanes, threads_per_cta=threads)
for i in range(2):
    r = ctx.sim_point("AWGN", snr, seed=0, point=0, frame0=i * frames, nframes=frames, decoding=dec, iterations=50, early_term=False)
    print(r, "algorithmic bytes per launch", frames * 50 * ctx.nnz * 32)
print(ctx.stats())
