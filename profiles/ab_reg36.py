"""Same-box A/B of library builds on a (3,6)-regular code (n = 1008, check degree 6: the fp64 sum-product body that is a call or inline in the
shared-memory kernel depending on B200_BP_CALL_FROM).  usage: python profiles/ab_reg36.py libA.so libB.so ..."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = "/tmp/reg36_n1008.txt"
if not os.path.exists(path):
    rng = np.random.default_rng(36)
    n, m = 1008, 504
    while True:   # configuration model without parallel edges
        sockets = np.repeat(np.arange(n), 3)
        rng.shuffle(sockets)
        rows = sockets.reshape(m, 6)
        if all(len(set(r)) == 6 for r in rows):
            break
    with open(path, "w") as f:
        f.write("nc: %d\nmc: %d\nnct: %d\nmct: %d\nnnz: %d\npuncture [0]: \nshorten [0]: \n" % (n, m, n, m, 3 * n))
        for i, r in enumerate(rows):
            for c in sorted(r):
                f.write("%d %d\n" % (i, c))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
from libldpc_b200 import api
ctx = api.Context(%r, "", device=0)
n = 148 * 4 * 64
out = []
for dec, x, et in (("BP", 1.0, False), ("BP", 2.5, True), ("BP_MS", 1.0, False)):
    ctx.sim_point("AWGN", x, nframes=n, decoding=dec, iterations=50, early_term=et)
    r = min((ctx.sim_point("AWGN", x, nframes=n, decoding=dec, iterations=50, early_term=et) for _ in range(3)), key=lambda r: r["device_ms"])
    out.append("%%s %%s%%+.1f %%.3f Gb/s (fec %%d its %%d)" %% (dec, "ET" if et else "fixed", x, n * ctx.nct / r["device_ms"] / 1e6, r["fec"], r["iters"]))
st = ctx.stats()
print(" | ".join(out), "| shape", st["frames_per_cta"], st["threads_per_cta"], st["ctas"])
''' % (ROOT, path)
for rnd in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, LDPC_B200_LIB=os.path.abspath(lib), LDPC_B200_TUNE_CACHE="off")
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(os.path.basename(lib), "|", (r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
