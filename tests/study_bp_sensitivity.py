"""Sensitivity study (test infrastructure, CPU only; not collected by pytest): how accurate must the Jacobian correction
log((1+e^-|x+y|)/(1+e^-|x-y|)) of sum-product decoding (reference src/decoding/decoder.h:12-15) be for the posteriors to stay
within the 1e-4 relative parity bar after 50 flooding iterations?

Builds a copy of the C oracle whose correction term is perturbed (PERT_MODE=1: a deterministic pseudo-random absolute error of
size PERT_EPS; PERT_MODE=2: the correction evaluated in float) and compares its posteriors with the unperturbed oracle on the
seeded parity-test inputs and on AWGN frames around the waterfall of the 1152x1024 sample code.
Result of 2026-10-18 (profiles/r1/bp_accuracy.md): non-converged frames amplify a perturbation by up to ~1e9, so anything
coarser than ~1e-14 absolute breaks the bar -- no mixed-precision or table shortcut for the fp64 sum-product kernel.

usage: python tests/study_bp_sensitivity.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = os.path.join(tempfile.gettempdir(), "bp_sensitivity")
HELPER = r'''
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
static double pert_eps(void){ static double e=-1; if(e<0){const char*s=getenv("PERT_EPS"); e=s?atof(s):0;} return e; }
static int pert_mode(void){ static int m=-1; if(m<0){const char*s=getenv("PERT_MODE"); m=s?atoi(s):0;} return m; }
static const double EXPC[14] = {1.0, 1.0, 1.0/2, 1.0/6, 1.0/24, 1.0/120, 1.0/720, 1.0/5040, 1.0/40320, 1.0/362880, 1.0/3628800, 1.0/39916800, 1.0/479001600, 1.0/6227020800.0};
static const double ATHC[17] = {1.0, 1.0/3, 1.0/5, 1.0/7, 1.0/9, 1.0/11, 1.0/13, 1.0/15, 1.0/17, 1.0/19, 1.0/21, 1.0/23, 1.0/25, 1.0/27, 1.0/29, 1.0/31, 1.0/33};
static double fast_exp_neg(double z)
{
    const double K0 = -1.4426950408889634, K1 = 6755399441055744.0, K2 = 0.6931471803691238, K3 = 1.9082149292705877e-10;
    double zc = fmin(z, 708.0), t = fma(zc, K0, K1), k = t - K1, r = fma(k, -K2, -zc), p = EXPC[13];
    uint64_t tb, pb;
    r = fma(k, -K3, r);
    for (int i = 12; i >= 0; --i) p = fma(p, r, EXPC[i]);
    memcpy(&tb, &t, 8); memcpy(&pb, &p, 8);
    pb = ((uint64_t)((uint32_t)(pb >> 32) + ((uint32_t)tb << 20)) << 32) | (pb & 0xffffffffu);
    memcpy(&p, &pb, 8);
    return p;
}
static double fast_log_ratio(double u, double v)
{
    double w = (u - v) / ((2.0 + u) + v), s = w * w, p = ATHC[16];
    for (int i = 15; i >= 0; --i) p = fma(p, s, ATHC[i]);
    return 2.0 * w * p;
}
static double pert_corr(double x, double y)
{
    double c = log((1 + exp(-fabs(x + y))) / (1 + exp(-fabs(x - y))));
    int m = pert_mode();
    if (m == 1) {
        union { double d; uint64_t u; } a; a.d = x * 1.618 + y; uint64_t h = a.u * 0x9E3779B97F4A7C15ull; h ^= h >> 29;
        return c + pert_eps() * (((double)(h & 0xFFFFF) / 524288.0) - 1.0);
    }
    if (m == 3) { /* the kernel's library-free evaluation (csrc/kernels.cuh bp_exp_neg / bp_log_ratio), same operations with fma() */
        return fast_log_ratio(fast_exp_neg(fabs(x + y)), fast_exp_neg(fabs(x - y)));
    }
    if (m == 2) {
        float a = fabsf((float)(x + y)), b = fabsf((float)(x - y));
        return (double)(logf((1.0f + expf(-a)) / (1.0f + expf(-b))));
    }
    return c;
}
'''


def build_perturbed():
    os.makedirs(WORK, exist_ok=True)
    src = open(os.path.join(ROOT, "oracle", "ldpc_oracle.c")).read()
    exact = "log((1 + exp(-fabs(x + y))) / (1 + exp(-fabs(x - y))))"
    assert src.count(exact) == 1
    src = src.replace(exact, "pert_corr(x, y)")
    i = src.index("static double f_jacobian")
    open(os.path.join(WORK, "o.c"), "w").write(src[:i] + HELPER + src[i:])
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "oracle"),
                    "-o", os.path.join(WORK, "liboracle_pert.so"), os.path.join(WORK, "o.c"), "-lm"], check=True)


def child(path):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    O.LIB = os.path.join(WORK, "liboracle_pert.so")
    code = O.Code(os.path.join(ROOT, "codes", "ref_h_n1152_m1024.txt"))
    rng = np.random.default_rng(1234)
    llr = rng.normal(1.0, 1.6, size=(96, code.nc))
    llr[:, code.puncture] = 0.0
    llr[3, :] = np.round(llr[3, :])
    llr[4, ::7] = -0.0
    llr[5, :] = np.where(rng.random(code.nc) < 0.2, -1.5, 1.5)
    llr[6, :200] = 99999.9
    sets = {"seeded": llr}
    for x in (-6.0, -5.0, -4.5):
        s2 = 10 ** (-x / 10)
        l = 2 * (1 + np.sqrt(s2) * rng.normal(size=(200, code.nc))) / s2
        l[:, code.puncture] = 0.0
        sets["awgn%g" % x] = l
    out = {}
    for k, l in sets.items():
        ro, rc, ri = code.decode(l, 50, True, False)
        out[k + "_o"], out[k + "_c"], out[k + "_i"] = ro, rc, ri
    np.savez(path, **out)


def run(mode, eps, path):
    subprocess.run([sys.executable, os.path.abspath(__file__), "child", path], check=True, env=dict(os.environ, PERT_MODE=str(mode), PERT_EPS=str(eps)))
    return np.load(path)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
        sys.exit(0)
    build_perturbed()
    base = run(0, 0, os.path.join(WORK, "base.npz"))
    print("| perturbation of the correction term | input set | max relative posterior error | frames over 1e-4 | identical decisions |")
    print("|---|---|---|---|---|")
    for mode, eps in ((3, 0), (1, 1e-15), (1, 1e-13), (1, 1e-11), (1, 1e-9), (1, 1e-7), (2, 0)):
        r = run(mode, eps, os.path.join(WORK, "p.npz"))
        for k in sorted(f for f in base.files if f.endswith("_o")):
            a, b = base[k], r[k]
            rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-9)
            same = (base[k[:-2] + "_c"] == r[k[:-2] + "_c"]).mean()
            print("| %s | %s | %.2e | %d / %d | %.6f |" % ("float evaluation" if mode == 2 else "kernel evaluation (library-free fp64)" if mode == 3 else "abs %.0e" % eps, k[:-2], rel.max(), (rel.max(1) > 1e-4).sum(), len(a), same))
