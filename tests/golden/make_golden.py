"""Generates the committed golden vectors from the UNMODIFIED reference (container only).

Runs oracle/_ref/dump_ref (the reference's own channel_* + ldpc_decoder classes, built by
oracle/Makefile from /root/reference) and oracle/_ref/ldpcsim_ref, and stores

  tests/golden/decode_cases.npz   per case: cw, llr_in, llr_out, co, iters  (a few frames each)
  tests/golden/curves.json        FER/BER points of the reference CLI (for binomial-CI checks)
  tests/golden/meta.json          code parameters, rank, compiler

Usage:  python tests/golden/make_golden.py          (needs /root/reference; ~2 min)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

H = os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt")
G = os.path.join(ROOT, "codes", "ref_g_k128_n1152.txt")
OUT = os.path.dirname(os.path.abspath(__file__))

# (name, channel, decoding, iterations, early_term, x, seed, frames, use_G)
CASES = [
    ("awgn_ms_et_m60", "AWGN", "BP_MS", 50, 1, -6.0, 3, 6, 0),
    ("awgn_ms_et_m45", "AWGN", "BP_MS", 50, 1, -4.5, 3, 8, 0),
    ("awgn_ms_et_m30", "AWGN", "BP_MS", 50, 1, -3.0, 3, 6, 0),
    ("awgn_ms_et_p20", "AWGN", "BP_MS", 50, 1, 2.0, 4, 4, 0),
    ("awgn_ms_noet5_m45", "AWGN", "BP_MS", 5, 0, -4.5, 5, 4, 0),
    ("awgn_ms_it1_m45", "AWGN", "BP_MS", 1, 1, -4.5, 6, 4, 0),
    ("awgn_ms_noet50_m50", "AWGN", "BP_MS", 50, 0, -5.0, 7, 4, 0),
    ("awgn_bp_et_m60", "AWGN", "BP", 50, 1, -6.0, 3, 6, 0),
    ("awgn_bp_et_m50", "AWGN", "BP", 50, 1, -5.0, 3, 8, 0),
    ("awgn_bp_et_m30", "AWGN", "BP", 50, 1, -3.0, 3, 6, 0),
    ("awgn_bp_noet2_m45", "AWGN", "BP", 2, 0, -4.5, 5, 4, 0),
    ("awgn_bp_noet50_m50", "AWGN", "BP", 50, 0, -5.0, 7, 4, 0),
    ("awgn_ms_G_m40", "AWGN", "BP_MS", 50, 1, -4.0, 8, 6, 1),
    ("awgn_bp_G_m45", "AWGN", "BP", 50, 1, -4.5, 8, 4, 1),
    ("bsc_ms_022", "BSC", "BP_MS", 50, 1, 0.22, 5, 8, 0),
    ("bsc_ms_010_G", "BSC", "BP_MS", 50, 1, 0.10, 5, 4, 1),
    ("bsc_bp_020", "BSC", "BP", 50, 1, 0.20, 5, 6, 0),
    ("bec_et_095", "BEC", "BP", 50, 1, 0.95, 5, 8, 0),
    ("bec_et_090_G", "BEC", "BP", 50, 1, 0.90, 5, 8, 1),
    ("bec_noet_080", "BEC", "BP", 10, 0, 0.80, 5, 4, 0),
]

# (name, extra CLI args) -> reference CLI curves, all with -t 8 -s 0
CURVES = [
    ("awgn_bp", ["-7", "-3.5", "0.5", "--decoding", "BP", "--frame-error-count", "150", "--max-frames", "60000"]),
    ("awgn_ms", ["-6", "-3.5", "0.5", "--decoding", "BP_MS", "--frame-error-count", "150", "--max-frames", "60000"]),
    ("bsc_ms", ["0.12", "0.24", "0.02", "--channel", "BSC", "--decoding", "BP_MS", "--frame-error-count", "150",
                "--max-frames", "60000"]),
    ("bsc_bp", ["0.14", "0.24", "0.02", "--channel", "BSC", "--decoding", "BP", "--frame-error-count", "150",
                "--max-frames", "60000"]),
]


def main():
    if not O.ref_available():
        O.build()
    store = {}
    for (name, ch, dec, it, et, x, seed, n, useg) in CASES:
        r = O.ref_sim_dump(H, ch, dec, it, et, x, seed, n, g=G if useg else None)
        for k, v in r.items():
            store[f"{name}/{k}"] = v
        store[f"{name}/cfg"] = np.array([it, et, seed, n, useg], dtype=np.int64)
        store[f"{name}/x"] = np.array([x])
        store[f"{name}/names"] = np.array([ch, dec])
        print(name, "iters", r["iters"])
    np.savez_compressed(os.path.join(OUT, "decode_cases.npz"), **store)

    curves = {}
    for (name, args) in CURVES:
        with tempfile.TemporaryDirectory() as td:
            res = os.path.join(td, "res.txt")
            subprocess.run([os.path.join(O.REF_DIR, "ldpcsim_ref"), H, res] + args + ["-t", "8", "-s", "0"],
                           check=True, stdout=subprocess.DEVNULL)
            pts = []
            for line in open(res).read().splitlines()[1:]:
                if not line.strip():
                    continue
                x, fer, ber, frames, avg_iter, _t = line.split()
                frames = int(frames)
                pts.append(dict(x=float(x), fer=float(fer), ber=float(ber), frames=frames,
                                fec=int(round(float(fer) * frames)), avg_iter=float(avg_iter)))
            curves[name] = dict(args=args, points=pts)
            print(name, [(p["x"], p["fer"]) for p in pts])
    json.dump(curves, open(os.path.join(OUT, "curves.json"), "w"), indent=1)

    c = O.Code(H)
    g = O.Code(G, matrix_only=True)
    meta = dict(h=os.path.basename(H), g=os.path.basename(G), nc=c.nc, mc=c.mc, nnz=c.nnz, nct=c.nct, mct=c.mct,
                kct=c.kct, kc=c.kc, max_degree=c.max_degree, rank=c.rank(), g_rows=g.mc, g_cols=g.nc, g_nnz=g.nnz,
                compiler=subprocess.run(["/usr/bin/g++", "--version"], capture_output=True, text=True).stdout.splitlines()[0],
                ldpctest_ref_rank=1021,
                note="vectors produced by the unmodified reference via oracle/_ref/dump_ref and ldpcsim_ref")
    json.dump(meta, open(os.path.join(OUT, "meta.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
