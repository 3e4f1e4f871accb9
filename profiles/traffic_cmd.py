"""One launch of a bench.py configuration's kernel inside an NVTX range, for the DRAM-traffic capture:
  ncu --nvtx --nvtx-include "measure/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv \
      python profiles/traffic_cmd.py <config>
Same code files, frame counts and parameters as bench.py (`headline` and the entries of `configs`).  LDPC_B200_PAIR=1 presets the
two-CTAs-per-SM shape for the shared-memory configurations (the timed trial is meaningless under the profiler)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "codes"))
import torch  # noqa: E402  (NVTX ranges)
import gen_codes  # noqa: E402
from libldpc_b200 import api  # noqa: E402

H = os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt")
CFG = {  # name: (code, channel, x, decoding, early_term, frames)
    "headline": ("h", "AWGN", -4.5, "BP_MS", False, 148 * 4 * 512),
    "C1_bp_fixed50": ("h", "AWGN", -4.5, "BP", False, 148 * 4 * 64),
    "C1_et_sweep": ("h", "AWGN", 2.0, "BP", True, 148 * 4 * 256),
    "C2_et": ("h", "AWGN", -4.5, "BP_MS", True, 148 * 4 * 256),
    "C5_bsc": ("h", "BSC", 0.08, "BP_MS", True, 1 << 22),
    "C5_bec": ("h", "BEC", 0.70, "BP", True, 1 << 24),
    "C3_bg1_ms": ("bg1", "AWGN", -0.5, "BP_MS", False, 148 * 32),
    "C4_dvbs2_bp_noet": ("dvbs2", "AWGN", 1.0, "BP", False, 148 * 16),
}
name = sys.argv[1]
code, ch, x, dec, et, frames = CFG[name]
path = H if code == "h" else gen_codes.ensure()[code]
ctx = api.Context(path, "", device=0)
if code != "h":
    ctx.prepare(dec, 50, et)   # the shape trial of a long sweep (kernels outside the NVTX range are not profiled)
ctx.sim_point(ch, x, seed=1, point=0, frame0=0, nframes=min(frames, 1 << 16), decoding=dec, iterations=50, early_term=et)   # warm-up / shape trial
ctx.stats(reset=True)
torch.cuda.nvtx.range_push("measure")
r = ctx.sim_point(ch, x, seed=2, point=0, frame0=0, nframes=frames, decoding=dec, iterations=50, early_term=et)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print(name, r, ctx.stats())
