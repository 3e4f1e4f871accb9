"""Global-residency tuning sweep on a large code: (precision, frames/CTA, threads/CTA) -> Gb/s, edge-it/s.
usage: python profiles/sweep_large.py codefile [decoding] [frames] [snr]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api  # noqa: E402

code = sys.argv[1]
dec = sys.argv[2] if len(sys.argv) > 2 else "BP_MS"
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 2400
snr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
ctx = api.Context(code, "", device=0)
for prec, pname, vec in ((api.F64, "f64", 2), (api.F32, "f32", 4)):
    for lanes in (1, 2, 4):
        for threads in (256, 512, 1024):
            if dec == "BP" and threads > 512:
                continue
            try:
                ctx.set_tuning(precision=prec, residency=api.GLOBAL, frames_per_cta=lanes * vec, threads_per_cta=threads, ctas=0)
                ctx.sim_point("AWGN", snr, nframes=600, decoding=dec, iterations=50, early_term=False)
                ctx.stats(reset=True)
                r = ctx.sim_point("AWGN", snr, nframes=frames, decoding=dec, iterations=50, early_term=False)
                st = ctx.stats()
                ms = r["device_ms"]
                row = dict(prec=pname, lanes=lanes, fpc=lanes * vec, threads=threads, ctas=st["ctas"], ms=round(ms, 3),
                           gbps=round(frames * ctx.nct / (ms * 1e-3) / 1e9, 4), gedge_it_s=round(frames * 50 * ctx.nnz / (ms * 1e-3) / 1e9, 2), fec=r["fec"])
            except RuntimeError as e:
                row = dict(prec=pname, lanes=lanes, threads=threads, error=str(e)[:80])
            print(json.dumps(row), flush=True)
