"""Which kernel shape do the automatic policies pick?  (LDPC_B200_DEBUG=1 prints the trial; run on a GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libldpc_b200 import api
ctx = api.Context('codes/ref_h_n1152_m1024.txt', '', device=0)
n = 148 * 4 * 64
_, llr = ctx.channel("AWGN", -4.5, 5, 0, 0, n)
out, hard, its = ctx.decode_batch(llr, "BP_MS", 50, False, want_llr=False)       # a decode-only user: the trial runs here
print("decode_batch only:", {k: ctx.stats()[k] for k in ("frames_per_cta", "threads_per_cta", "ctas")})
for dec in ("BP_MS", "BP"):
    for prec in (api.F64, api.F32):
        ctx.set_tuning(precision=prec)
        r = ctx.sim_point("AWGN", -4.5, nframes=n, decoding=dec, iterations=50, early_term=False)
        r = ctx.sim_point("AWGN", -4.5, nframes=n, decoding=dec, iterations=50, early_term=False)
        st = ctx.stats()
        print(dec, "f32" if prec else "f64", st["frames_per_cta"], st["threads_per_cta"], st["ctas"], "%.3f Gb/s" % (n * 1024 / r["device_ms"] / 1e6))
