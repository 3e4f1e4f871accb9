"""Where does the host-buffer decode path spend its time?  Device-resident decode vs pinned-host decode_batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from libldpc_b200 import api

ctx = api.Context(os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt"), "", device=0)
nb = 148 * 4 * 96
_, gen = ctx.channel("AWGN", -4.5, 5, 0, 0, nb)
d_llr = torch.from_numpy(gen).cuda()
d_hard = torch.empty((nb, 1152), dtype=torch.uint8, device="cuda")
d_its = torch.empty(nb, dtype=torch.int32, device="cuda")
d_out = torch.empty((nb, 1152), dtype=torch.float64, device="cuda")
s = torch.cuda.Stream()
for name, po, ph in (("dev: hard+iters", 0, d_hard.data_ptr()), ("dev: llr_out+hard+iters", d_out.data_ptr(), d_hard.data_ptr()), ("dev: iters only", 0, 0)):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.decode_batch_device(d_llr.data_ptr(), nb, po, ph, d_its.data_ptr(), s.cuda_stream, "BP_MS", 50, False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(name, "%.3f Gb/s" % (nb * 1024 / dt / 1e9))
pin_in = torch.empty((nb, 1152), dtype=torch.float64, pin_memory=True); pin_in.numpy()[:] = gen
pin_hard = torch.empty((nb, 1152), dtype=torch.uint8, pin_memory=True)
pin_its = torch.empty(nb, dtype=torch.int32, pin_memory=True)
for rep in range(3):
    t0 = time.perf_counter()
    ctx.decode_batch(pin_in.numpy(), "BP_MS", 50, False, want_llr=False, hard=pin_hard.numpy(), its=pin_its.numpy())
    dt = time.perf_counter() - t0
print("pinned host: hard+iters %.3f Gb/s" % (nb * 1024 / dt / 1e9))
t0 = time.perf_counter(); d2 = pin_in.cuda(non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plain pinned H2D %.1f GB/s" % (pin_in.numel() * 8 / dt / 1e9))
r = ctx.sim_point("AWGN", -4.5, nframes=nb, decoding="BP_MS", iterations=50, early_term=False)
print("fused sim_point %.3f Gb/s" % (nb * 1024 / (r["device_ms"] * 1e-3) / 1e9))
