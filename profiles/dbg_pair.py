import os, sys
sys.path.insert(0, '.')
from libldpc_b200 import api
ctx = api.Context('codes/ref_h_n1152_m1024.txt', '', device=0)
n = 148 * 4 * 512
for kw in (dict(frames_per_cta=2, threads_per_cta=256, idx16=2, ctas=0), dict(frames_per_cta=2, threads_per_cta=256, idx16=2, ctas=296),
           dict(frames_per_cta=2, threads_per_cta=256, idx16=2, ctas=148), dict(frames_per_cta=4, threads_per_cta=512, idx16=1, ctas=0),
           dict(frames_per_cta=0, threads_per_cta=0, idx16=0, ctas=0)):
    ctx.set_tuning(precision=api.F64, residency=api.SMEM, **kw)
    for et in (False, True):
        ctx.sim_point("AWGN", -4.5, nframes=20000, decoding="BP_MS", iterations=50, early_term=et)
        best = min(ctx.sim_point("AWGN", -4.5, nframes=n, decoding="BP_MS", iterations=50, early_term=et)["device_ms"] for _ in range(3))
        st = ctx.stats()
        print(kw, et, st['frames_per_cta'], st['threads_per_cta'], st['ctas'], st['smem_bytes'], "ms %.3f Gb/s %.3f" % (best, n * 1024 / best / 1e6), flush=True)
