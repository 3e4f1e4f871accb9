"""Same-box A/B of library builds on the sum-product workloads: h.txt (shared-memory residency) fixed 50 iterations and an
early-termination point, and the DVB-S2-shaped code (global residency).  usage: python profiles/ab_bp.py libA.so libB.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "codes"))
import gen_codes
from libldpc_b200 import api
out = []
ctx = api.Context(os.path.join(%r, "codes", "ref_h_n1152_m1024.txt"), "", device=0)
n = 148 * 4 * 64
for x, et in ((-4.5, False), (2.0, True)):
    ctx.sim_point("AWGN", x, nframes=n, decoding="BP", iterations=50, early_term=et)
    r = min((ctx.sim_point("AWGN", x, nframes=n, decoding="BP", iterations=50, early_term=et) for _ in range(3)), key=lambda r: r["device_ms"])
    out.append("h.txt %%s%%+.1f %%.3f Gb/s (fec %%d its %%d)" %% ("ET" if et else "fixed", x, n * 1024 / r["device_ms"] / 1e6, r["fec"], r["iters"]))
ctx.close()
big = gen_codes.ensure()
ctx = api.Context(big["dvbs2"], "", device=0)
ctx.sim_point("AWGN", 1.0, nframes=20480, decoding="BP", iterations=50, early_term=False)
r = min((ctx.sim_point("AWGN", 1.0, nframes=2048, decoding="BP", iterations=50, early_term=False) for _ in range(2)), key=lambda r: r["device_ms"])
out.append("dvbs2 %%.1f G edge-it/s" %% (2048 * 50 * ctx.nnz / r["device_ms"] / 1e6))
print(" | ".join(out))
''' % (ROOT, ROOT, ROOT)
for rnd in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, LDPC_B200_LIB=os.path.abspath(lib), LDPC_B200_TUNE_CACHE="off")
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(os.path.basename(lib), "|", (r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
