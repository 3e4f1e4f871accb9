"""A/B timing of tuning variants on the bench workload (h.txt, AWGN -4.5 dB, 50 fixed iterations).
usage: python profiles/ab.py [decoding] [frames]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api  # noqa: E402

dec = sys.argv[1] if len(sys.argv) > 1 else "BP_MS"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 4 * 128
ctx = api.Context(os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt"), "", device=0)
for prec, pname, vec in ((api.F64, "f64", 2), (api.F32, "f32", 4)):
    for lanes in [int(t) for t in os.environ.get('AB_LANES', '1,2').split(',')]:
        for threads in [int(t) for t in os.environ.get('AB_THREADS', '256,512').split(',')]:
            for tmem in (1, 0):
                for et in ((False, True) if os.environ.get('AB_ET', '1') == '1' else (False,)):
                    try:
                        ctx.set_tuning(precision=prec, residency=api.SMEM, frames_per_cta=lanes * vec, threads_per_cta=threads, ctas=0, tmem=tmem, idx16=int(os.environ.get('AB_IDX16', '1')))
                        ctx.sim_point("AWGN", -4.5, nframes=2000, decoding=dec, iterations=50, early_term=et)
                        best = None
                        for rep in range(3):
                            r = ctx.sim_point("AWGN", -4.5, nframes=frames, decoding=dec, iterations=50, early_term=et)
                            best = r["device_ms"] if best is None else min(best, r["device_ms"])
                        st = ctx.stats()
                        row = dict(prec=pname, lanes=lanes, threads=threads, tmem="off" if tmem else "on", et=et, ctas=st['ctas'], smem=st['smem_bytes'], ms=round(best, 3),
                                   gbps=round(frames * ctx.nct / (best * 1e-3) / 1e9, 3), fec=r["fec"], iters=r["iters"])
                    except RuntimeError as e:
                        row = dict(prec=pname, lanes=lanes, threads=threads, tmem=tmem, error=str(e)[:80])
                    print(json.dumps(row), flush=True)
