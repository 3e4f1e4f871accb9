"""Layered vs flooding min-sum on the BG1-shaped code (and h.txt): FER, average iterations, throughput per SNR point.
usage: python profiles/layered_table.py [frames_per_point]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "codes"))
import gen_codes
from libldpc_b200 import api
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
big = gen_codes.ensure()
for name, path, xs, n in (("BG1-shaped Z=384", big["bg1"], [-1.0, -0.8, -0.6, -0.4, -0.2], frames),
                          ("h.txt", os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt"), [-5.5, -5.0, -4.5, -4.0], frames * 20)):
    ctx = api.Context(path, "", device=0)
    print(f"\n### {name}: BP_MS fp64, -i 50, early termination, {n} frames per point ({len(ctx.layers())} layers)\n")
    cols = (("flooding BP_MS", api.FLOODING, "BP_MS", 0), ("layered BP_MS", api.LAYERED, "BP_MS", 0), ("layered BP_MS x0.75", api.LAYERED, "BP_MS", 48),
            ("flooding BP", api.FLOODING, "BP", 0), ("layered BP", api.LAYERED, "BP", 0))
    print("| Es/sigma^2 dB | " + " | ".join(f"{c[0]}: FER / avg it / frames/s" for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for x in xs:
        row = []
        for _, sched, dec, q in cols:
            nn = n if dec == "BP_MS" else max(n // 4, 1000)
            ctx.set_tuning(schedule=sched, layered_ms_scale64=q)
            ctx.sim_point("AWGN", x, seed=1, point=0, frame0=0, nframes=min(nn, 4096), decoding=dec, iterations=50, early_term=True)
            r = ctx.sim_point("AWGN", x, seed=1, point=0, frame0=0, nframes=nn, decoding=dec, iterations=50, early_term=True)
            row.append(f"{r['fec'] / nn:.2e} / {r['iters'] / nn:.1f} / {nn / r['device_ms'] * 1e3:.3g}")
        print(f"| {x:g} | " + " | ".join(row) + " |", flush=True)
    ctx.close()
