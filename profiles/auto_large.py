"""Default (autotuned) global-residency configuration on a large code: chosen tile shape and throughput.
usage: python profiles/auto_large.py codefile [decoding] [frames] [snr]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api  # noqa: E402

code = sys.argv[1]
dec = sys.argv[2] if len(sys.argv) > 2 else "BP_MS"
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 24000
snr = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
ctx = api.Context(code, "", device=0)
for prec, pname in ((api.F64, "f64"), (api.F32, "f32")):
    ctx.set_tuning(precision=prec, residency=api.AUTO, frames_per_cta=0, threads_per_cta=0, ctas=0)
    ctx.sim_point("AWGN", snr, nframes=frames, decoding=dec, iterations=50, early_term=False)   # includes the one-off autotune
    ctx.stats(reset=True)
    r = ctx.sim_point("AWGN", snr, nframes=frames, decoding=dec, iterations=50, early_term=False)
    st = ctx.stats()
    ms = r["device_ms"]
    print(json.dumps(dict(code=os.path.basename(code), prec=pname, dec=dec, frames=frames, fpc=st["frames_per_cta"], threads=st["threads_per_cta"],
                          ctas=st["ctas"], ms=round(ms, 2), gbps=round(frames * ctx.nct / (ms * 1e-3) / 1e9, 4),
                          gedge_it_s=round(frames * 50 * ctx.nnz / (ms * 1e-3) / 1e9, 1), fec=r["fec"])), flush=True)
