"""Generates the irregular-code golden vectors from the UNMODIFIED reference (container only).

The irregular code (tests/golden/irregular_h.txt, written here once and committed) is a small code that fits
shared memory and exercises every node-update body of the CUDA kernels: check degrees 2..13 (fixed-degree
bodies 2..8 and the generic path), variable degrees 0..~25 (bodies 0..16 and the chunked path), punctured and
shortened positions, an explicit zero-valued entry.  No reference-produced vector covered these degrees in
round 1 (the reference's own fixture h.txt has check degree 3-4, variable degree 1/2/15).

  tests/golden/irregular_cases.npz   per case: llr_in, llr_out, co, iters — outputs of
                                     oracle/_ref/dump_ref decode (the reference's ldpc_decoder::decode,
                                     /root/reference/src/decoding/decoder.cpp:11-78) on seeded inputs

Usage:  python tests/golden/make_irregular_golden.py      (needs /root/reference; seconds)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
H_IRR = os.path.join(OUT, "irregular_h.txt")

# (name, decoding, iterations, early_term, input kind)
CASES = [
    ("ms_et40", "BP_MS", 40, 1, "gauss"),
    ("ms_noet9", "BP_MS", 9, 0, "gauss"),
    ("ms_et50_hard", "BP_MS", 50, 1, "gauss_hard"),
    ("ms_bsc", "BP_MS", 50, 1, "bsc"),
    ("bp_it2", "BP", 2, 0, "gauss"),
    ("bp_et50", "BP", 50, 1, "gauss"),
    ("bp_noet50", "BP", 50, 0, "gauss_hard"),
    ("bp_bsc", "BP", 25, 1, "bsc"),
    ("ms_et50_easy", "BP_MS", 50, 1, "gauss_easy"),
    ("bp_et50_easy", "BP", 50, 1, "gauss_easy"),
]


def write_code(path=H_IRR):
    rng = np.random.default_rng(2024)
    nc, mc = 420, 240
    w = 1.0 / (1 + np.arange(nc)) ** 0.7          # skewed column popularity -> a wide spread of variable degrees
    w[-3:] = 0                                    # three variables without any edge (degree 0)
    w /= w.sum()
    edges = []
    for r in range(mc):
        deg = 2 + (r % 12)
        cols = rng.choice(nc, size=deg, replace=False, p=w)
        edges += [(r, int(c)) for c in cols]
    edges.append((mc - 1, nc - 4))                # make sure the last-but-three column exists with degree 1
    edges = sorted(set(edges), key=lambda e: (e[0], rng.random()))   # rows grouped, columns in random file order
    with open(path, "w") as f:
        f.write("puncture [3]: 0 5 9 \nshorten [2]: 17 33 \n")
        f.write("\n".join(f"{r} {c}" for r, c in edges) + "\n")
        f.write(f"{mc - 1} {nc - 1} 0\n")         # explicit zero-valued entry pins nc (stored as 1 by the reference: sparse.h:124)
    return path


def inputs(kind, oc, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "gauss":
        llr = rng.normal(1.4, 1.9, size=(n, oc.nc))
    elif kind == "gauss_easy":
        llr = rng.normal(2.6, 1.7, size=(n, oc.nc))
    elif kind == "gauss_hard":
        llr = rng.normal(0.55, 1.6, size=(n, oc.nc))
    else:                                         # two-valued LLRs: exact ties and exact zeros are common
        delta = np.log((1 - 0.06) / 0.06)
        llr = np.where(rng.random((n, oc.nc)) < 0.06, -delta, delta)
    llr[:, oc.puncture] = 0.0
    llr[:, oc.shorten] = 99999.9 if kind != "bsc" else np.log((1 - 0.06) / 0.06)
    if kind != "bsc":
        llr[1, ::5] = -0.0                        # sign of zero matters (std::signbit)
        llr[2, ::7] = 0.0
    return llr


def main():
    if not O.ref_available():
        O.build()
    if not os.path.exists(H_IRR):
        write_code()
    oc = O.Code(H_IRR)
    store = {}
    for i, (name, dec, it, et, kind) in enumerate(CASES):
        llr = inputs(kind, oc, 6, 100 + i)
        r = O.ref_decode(H_IRR, dec, it, bool(et), llr)
        store[f"{name}/llr_in"] = llr
        store[f"{name}/llr_out"] = r["llr_out"]
        store[f"{name}/co"] = r["co"]
        store[f"{name}/iters"] = r["iters"]
        store[f"{name}/cfg"] = np.array([it, et], dtype=np.int64)
        store[f"{name}/names"] = np.array([dec, kind])
        print(name, "iters", r["iters"])
    np.savez_compressed(os.path.join(OUT, "irregular_cases.npz"), **store)


if __name__ == "__main__":
    main()
