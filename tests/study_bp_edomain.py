"""Accuracy study (test infrastructure, CPU only; not collected by pytest) of the sum-product check-node evaluation the fp64 CUDA
kernel uses since round 2: instead of 3(d-2) pairwise box-plus operations (reference src/decoding/decoder.h:12-15 through the
forward/backward recursion of src/decoding/decoder.cpp:30-44; two exp, a division and a log each) the check works on
E = e^-|x|, where the box-plus magnitude is (Ex + Ey) / (1 + Ex Ey): d exponentials in, fraction-valued forward/backward
products (no division), d logarithms out.  Mathematically the same function; this script emulates the kernel's operation
sequence with fma() inside a copy of the C oracle and compares the posteriors with the unmodified oracle (= the reference).

usage: python tests/study_bp_edomain.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = os.path.join(tempfile.gettempdir(), "bp_edomain")
HELPER = open(os.path.join(ROOT, "tests", "bp_edomain_emul.h")).read()


def build():
    os.makedirs(WORK, exist_ok=True)
    src = open(os.path.join(ROOT, "oracle", "ldpc_oracle.c")).read()
    hook = "if (cw < 2) continue; /* degree 0/1 checks index out of range in the reference (UB); skipped */"
    assert src.count(hook) == 1
    src = src.replace(hook, hook + "\n            if (!minsum && getenv(\"EDOMAIN\")) { double xin[128], xo[128]; for (int j = 0; j < cw; ++j) xin[j] = v2c[cn[j]];"
                      " edomain_check(xin, cw, xo); for (int j = 0; j < cw; ++j) c2v[cn[j]] = xo[j]; continue; }")
    i = src.index("int orc_is_codeword")
    open(os.path.join(WORK, "o.c"), "w").write(src[:i] + HELPER + src[i:])
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "oracle"),
                    "-o", os.path.join(WORK, "liboracle_e.so"), os.path.join(WORK, "o.c"), "-lm"], check=True)


def child(path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "codes"))
    from oracle import oracle as O
    O.LIB = os.path.join(WORK, "liboracle_e.so")
    out = {}
    code = O.Code(os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt"))
    rng = np.random.default_rng(1234)
    llr = rng.normal(1.0, 1.6, size=(96, code.nc))
    llr[:, code.puncture] = 0.0
    llr[3, :] = np.round(llr[3, :])
    llr[4, ::7] = -0.0
    llr[5, :] = np.where(rng.random(code.nc) < 0.2, -1.5, 1.5)
    llr[6, :200] = 99999.9
    llr[7, :] *= 300.0
    llr[8, ::3] = 99999.9
    llr[9, ::2] *= 1e-9
    sets = {"seeded": (code, llr, True)}
    for x in (-6.0, -5.0, -4.5, 2.0):
        s2 = 10 ** (-x / 10)
        l = 2 * (1 + np.sqrt(s2) * rng.normal(size=(200, code.nc))) / s2
        l[:, code.puncture] = 0.0
        sets["awgn%g" % x] = (code, l, True)
        sets["awgn%g fixed50" % x] = (code, l[:60], False)
    if os.environ.get("LARGE"):
        import gen_codes
        big = gen_codes.ensure()
        for name, x in (("dvbs2", 1.0), ("bg1", -0.5)):
            c = O.Code(big[name])
            s2 = 10 ** (-x / 10)
            l = 2 * (1 + np.sqrt(s2) * rng.normal(size=(6, c.nc))) / s2
            l[:, c.puncture] = 0.0
            sets[name + " fixed50"] = (c, l, False)
    for k, (c, l, et) in sets.items():
        ro, rc, ri = c.decode(l, 50, et, False)
        out[k + "_o"], out[k + "_c"], out[k + "_i"] = ro, rc, ri
    np.savez(path, **out)


def run(env, path):
    subprocess.run([sys.executable, os.path.abspath(__file__), "child", path], check=True, env=dict(os.environ, **env))
    return np.load(path)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
        sys.exit(0)
    build()
    base = run({}, os.path.join(WORK, "base.npz"))
    r = run({"EDOMAIN": "1"}, os.path.join(WORK, "e.npz"))
    print("| input set | max abs posterior | max relative posterior error | frames over 1e-4 | identical decisions | identical iteration counts |")
    print("|---|---|---|---|---|---|")
    for k in sorted(f for f in base.files if f.endswith("_o")):
        a, b = base[k], r[k]
        rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-9)
        rel = np.where(np.isfinite(rel), rel, np.inf if not np.array_equal(np.isnan(a), np.isnan(b)) else 0.0)
        same = (base[k[:-2] + "_c"] == r[k[:-2] + "_c"]).mean()
        its = (base[k[:-2] + "_i"] == r[k[:-2] + "_i"]).mean()
        print("| %s | %.3g | %.2e | %d / %d | %.6f | %.4f |" % (k[:-2], np.nanmax(np.abs(a)), rel.max(), (rel.max(1) > 1e-4).sum(), len(a), same, its))
