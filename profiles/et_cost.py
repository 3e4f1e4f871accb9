"""Cost model of the early-termination path on h.txt (f64 min-sum): per-iteration cost with and without the early-termination
instantiation at an SNR where nothing converges, and the per-frame cost at SNRs where frames converge at once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libldpc_b200 import api
ctx = api.Context('codes/ref_h_n1152_m1024.txt', '', device=0)
n = 148 * 4 * 256
ctx.sim_point("AWGN", -4.5, nframes=n, decoding="BP_MS", iterations=50, early_term=False)   # shape trial
for x in (-9.0, -6.0, -5.0, -4.5, -4.0, -3.0, 0.0, 3.0):
    row = []
    for et in (False, True):
        ctx.sim_point("AWGN", x, nframes=n, decoding="BP_MS", iterations=50, early_term=et)
        r = min((ctx.sim_point("AWGN", x, nframes=n, decoding="BP_MS", iterations=50, early_term=et) for _ in range(2)), key=lambda r: r["device_ms"])
        row.append(r)
    f, e = row
    it = e["iters"] / n
    print("x %5.1f  fixed %.2f ns/frame-it | ET: avg it %5.2f  %.1f ns/frame = %.2f ns/frame-it  FER %.4f  | frames/s %.3g  Gb/s %.2f" % (
        x, f["device_ms"] * 1e6 / (n * 50), it, e["device_ms"] * 1e6 / n, e["device_ms"] * 1e6 / (n * it), e["fec"] / n, n / e["device_ms"] * 1e3, n * 1024 / e["device_ms"] / 1e6), flush=True)
