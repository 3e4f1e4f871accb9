"""Stage timing of the refill events (phase-timing debug build): B200_PHASE_TIMING=1 python -m libldpc_b200.build, then
LDPC_B200_LIB=libldpc_b200/libldpc_pt.so LDPC_B200_PAIR=1 python profiles/et_phase.py [snr] [decoding] [channel]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libldpc_b200 import api
x = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
dec = sys.argv[2] if len(sys.argv) > 2 else "BP_MS"
ch = sys.argv[3] if len(sys.argv) > 3 else "AWGN"
ctx = api.Context('codes/ref_h_n1152_m1024.txt', '', device=0)
n = 148 * 2 * 2 * 200
for et in (True,):
    r = ctx.sim_point(ch, x, nframes=n, decoding=dec, iterations=50, early_term=et)
    print("x", x, dec, ch, "et", et, "ns/frame %.2f" % (r["device_ms"] * 1e6 / n), "avg it %.2f" % (r["iters"] / n), ctx.stats(), flush=True)
