"""Same-box A/B of two builds of the library on the bench workload (h.txt, AWGN): fixed 50 iterations and early termination.
usage: python profiles/ab_lib.py libA.so libB.so ...   (each runs in its own process; alternating, 2 rounds)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
from libldpc_b200 import api
ctx = api.Context(os.path.join(%r, "codes", "ref_h_n1152_m1024.txt"), "", device=0)
n = 148 * 4 * 256
ctx.sim_point("AWGN", -4.5, nframes=n, decoding="BP_MS", iterations=50, early_term=False)
out = []
for x, et in ((-4.5, False), (-4.5, True), (0.0, True), (3.0, True)):
    ctx.sim_point("AWGN", x, nframes=n, decoding="BP_MS", iterations=50, early_term=et)
    r = min((ctx.sim_point("AWGN", x, nframes=n, decoding="BP_MS", iterations=50, early_term=et) for _ in range(3)), key=lambda r: r["device_ms"])
    out.append("%%s%%+.1f: %%.2f ns/frame (it %%.2f)" %% ("ET" if et else "fixed", x, r["device_ms"] * 1e6 / n, r["iters"] / n))
print(" | ".join(out))
''' % (ROOT, ROOT)
for rnd in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, LDPC_B200_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(os.path.basename(lib), "|", (r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
