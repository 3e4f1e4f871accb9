"""Min-sum on h.txt: one CTA per SM (4 frames x 512 threads) against two (2 x 256, 16-bit tables) on early-termination workloads.
usage: LDPC_B200_PAIR=0|1 python profiles/shape_ab_ms.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libldpc_b200 import api
ctx = api.Context(os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt"), "", device=0)
out = []
for ch, x, et, n in (("AWGN", -4.5, False, 148 * 4 * 256), ("AWGN", -4.5, True, 148 * 4 * 256), ("AWGN", 0.0, True, 148 * 4 * 512), ("AWGN", 3.0, True, 148 * 4 * 1024),
                     ("BSC", 0.08, True, 1 << 21), ("BSC", 0.04, True, 1 << 21)):
    ctx.sim_point(ch, x, nframes=n, decoding="BP_MS", iterations=50, early_term=et)
    r = min((ctx.sim_point(ch, x, nframes=n, decoding="BP_MS", iterations=50, early_term=et) for _ in range(3)), key=lambda r: r["device_ms"])
    out.append("%s %g %s: %.2f ns/frame" % (ch, x, "ET" if et else "fixed", r["device_ms"] * 1e6 / n))
st = ctx.stats()
print("pair=%s (%d x %d x %d) | " % (os.environ.get("LDPC_B200_PAIR"), st["frames_per_cta"], st["threads_per_cta"], st["ctas"]) + " | ".join(out))
