"""Soak comparison (test infrastructure, GPU box; not collected by pytest): the fp64 sum-product CUDA path against the C oracle on many frames,
including scaled inputs that exercise the shifted exponentials / the exact path of the check node (csrc/kernels.cuh bp_check).
usage: python tests/soak_bp_gpu.py [frames_per_set]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "codes"))
from oracle import oracle as O  # noqa: E402
from libldpc_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
H = os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt")
IRR = os.path.join(ROOT, "tests", "golden", "irregular_h.txt")
print("| code | input set | early termination | frames | max relative posterior error | frames over 1e-4 | identical decisions | identical iteration counts |")
print("|---|---|---|---|---|---|---|---|")
for cname, path in (("h.txt", H), ("irregular (degrees up to 81, shortened positions)", IRR)):
    oc = O.Code(path)
    ctx = api.Context(path, "", device=0)
    rng = np.random.default_rng(99)
    sets = {}
    for x in (-6.0, -5.0, -4.5, -4.0, 0.0):
        s2 = 10 ** (-x / 10)
        l = 2 * (1 + np.sqrt(s2) * rng.normal(size=(n, oc.nc))) / s2
        sets["awgn %+.1f dB" % x] = l
    base = sets["awgn -4.5 dB"]
    for f in (8.0, 40.0, 300.0, 5000.0):
        sets["awgn -4.5 dB x %g" % f] = base * f
    sets["awgn -4.5 dB x 1e-6"] = base * 1e-6
    mix = base.copy()
    mix[:, ::5] *= 1000.0
    sets["awgn -4.5 dB, every 5th x 1000"] = mix
    for k, l in sets.items():
        l = l.copy()
        l[:, oc.puncture] = 0.0
        if len(oc.shorten):
            l[:, oc.shorten] = 99999.9
        for et in (True, False):
            ro, rc, ri = oc.decode(l, 50, et, False)
            out, hard, its = ctx.decode_batch(l, "BP", 50, et)
            rel = np.abs(out - ro) / np.maximum(np.abs(ro), 1e-9)
            rel = np.where(np.isfinite(ro) & np.isfinite(out), rel, np.where(out == ro, 0.0, np.inf))
            print("| %s | %s | %s | %d | %.2e | %d | %.6f | %.4f |" % (cname, k, et, n, rel.max(), (rel.max(1) > 1e-4).sum(), (hard == rc).mean(), (its == ri).mean()), flush=True)
    ctx.close()
