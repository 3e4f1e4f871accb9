"""Same-box A/B of library builds on the two large codes (global residency): fixed 50 iterations, min-sum and sum-product, fp64.
usage: python profiles/ab_large.py libA.so libB.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "codes"))
import gen_codes
from libldpc_b200 import api
big = gen_codes.ensure()
out = []
for name, x, n in (("bg1", -0.5, 4096), ("dvbs2", 1.0, 2048)):
    ctx = api.Context(big[name], "", device=0)
    for dec in ("BP_MS", "BP"):
        ctx.sim_point("AWGN", x, nframes=20480, decoding=dec, iterations=50, early_term=False)   # shape trial + warm-up
        r = min((ctx.sim_point("AWGN", x, nframes=n, decoding=dec, iterations=50, early_term=False) for _ in range(2)), key=lambda r: r["device_ms"])
        st = ctx.stats()
        out.append("%%s %%s %%.1f G edge-it/s (fpc %%d thr %%d ctas %%d)" %% (name, dec, n * 50 * ctx.nnz / r["device_ms"] / 1e6, st["frames_per_cta"], st["threads_per_cta"], st["ctas"]))
    ctx.close()
print(" | ".join(out))
''' % (ROOT, ROOT)
for lib in sys.argv[1:]:
    env = dict(os.environ, LDPC_B200_LIB=os.path.abspath(lib), LDPC_B200_TUNE_CACHE="off")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), "|", (r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
