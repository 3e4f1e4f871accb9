"""Short erasure-sweep workload for ncu captures of the bit-sliced kernel.  usage: python profiles/profile_bec.py [frames] [eps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 16
eps = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
ctx = api.Context(os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt"), "", device=0)
for i in range(3):
    print(ctx.sim_point("BEC", eps, seed=0, point=0, frame0=i * frames, nframes=frames, decoding="BP", iterations=50, early_term=True))
print(ctx.stats())
