"""Tuning sweep of the tile kernel on one code: (precision, frames/CTA, threads/CTA, CTAs) -> Gb/s.
usage: python profiles/sweep.py [codefile] [decoding] [frames] [snr]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api  # noqa: E402

code = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt")
dec = sys.argv[2] if len(sys.argv) > 2 else "BP_MS"
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 60000
snr = float(sys.argv[4]) if len(sys.argv) > 4 else -4.5
ctx = api.Context(code, "", device=0)
rows = []
for prec, pname, vec in ((api.F64, "f64", 2), (api.F32, "f32", 4)):
    for res in (api.SMEM, api.GLOBAL):
        for lanes in (1, 2, 4, 8):
            for threads in (256, 512, 1024):
                if dec == "BP" and threads > 512:
                    continue
                try:
                    ctx.set_tuning(precision=prec, residency=res, frames_per_cta=lanes * vec, threads_per_cta=threads, ctas=0)
                    ctx.sim_point("AWGN", snr, nframes=2000, decoding=dec, iterations=50, early_term=False)
                    best = None
                    for rep in range(2):
                        ctx.stats(reset=True)
                        r = ctx.sim_point("AWGN", snr, nframes=frames, decoding=dec, iterations=50, early_term=False)
                        best = r["device_ms"] if best is None else min(best, r["device_ms"])
                    st = ctx.stats()
                    gbps = frames * ctx.nct / (best * 1e-3) / 1e9
                    row = dict(prec=pname, res="smem" if res == api.SMEM else "global", lanes=lanes, fpc=lanes * vec, threads=threads,
                               ctas=st["ctas"], smem=st["smem_bytes"], ms=round(best, 3), gbps=round(gbps, 3), fec=r["fec"])
                except RuntimeError as e:
                    row = dict(prec=pname, res="smem" if res == api.SMEM else "global", lanes=lanes, threads=threads, error=str(e)[:80])
                rows.append(row)
                print(json.dumps(row), flush=True)
