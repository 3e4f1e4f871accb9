"""Minimax polynomials of the fp64 sum-product check node (csrc/kernels.cuh BP_EXP_C / BP_LOG_C) by the Remez exchange at 60 digits (mpmath):
e^r on |r| <= ln 2 / 2 (degree 11) and 2 atanh(w) / w as a function of s = w^2 on [0, 0.0405] (degree 7).  Prints the coefficients as
hexadecimal binary64 literals and the maximum error of the rounded polynomials.  usage: python profiles/gen_bp_tables.py"""
import mpmath as mp

mp.mp.dps = 60


def remez(f, a, b, n, iters=12, N=4000):
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * k / (n + 1)) for k in range(n + 2)][::-1]
    grid = [a + (b - a) * i / N for i in range(N + 1)]
    for _ in range(iters):
        A = mp.matrix(n + 2, n + 2)
        y = mp.matrix(n + 2, 1)
        for i, x in enumerate(xs):
            for j in range(n + 1):
                A[i, j] = x ** j
            A[i, n + 1] = (-1) ** i
            y[i] = f(x)
        sol = mp.lu_solve(A, y)
        c = [sol[j] for j in range(n + 1)]
        ev = [sum(c[j] * x ** j for j in range(n + 1)) - f(x) for x in grid]
        ext = [(grid[i], ev[i]) for i in range(N + 1) if (i == 0 or abs(ev[i]) >= abs(ev[i - 1])) and (i == N or abs(ev[i]) >= abs(ev[i + 1]))]
        sel = []
        for x, e in ext:  # alternating extrema, the largest of every run of equal signs
            if sel and (e > 0) == (sel[-1][1] > 0):
                if abs(e) > abs(sel[-1][1]):
                    sel[-1] = (x, e)
            else:
                sel.append((x, e))
        while len(sel) > n + 2:
            sel.pop(0) if abs(sel[0][1]) < abs(sel[-1][1]) else sel.pop()
        if len(sel) < n + 2:
            break
        xs = [x for x, _ in sel]
    return c


def report(name, f, a, b, n, scale=1):
    c = [float(scale * x) for x in remez(f, a, b, n)]
    N = 2000
    err = max(abs(sum(mp.mpf(c[j]) * (a + (b - a) * i / N) ** j for j in range(n + 1)) - scale * f(a + (b - a) * i / N)) for i in range(N + 1))
    print("%s: degree %d, max error with binary64 coefficients %.2e" % (name, n, err))
    print("    " + ", ".join(x.hex() for x in c))


if __name__ == "__main__":
    h = mp.log(2) / 2 * mp.mpf("1.0001")
    report("BP_EXP_C  e^r, |r| <= ln2/2", mp.exp, -h, h, 11)
    report("BP_LOG_C  2 atanh(sqrt s)/sqrt s, s in [0, 0.0405]", lambda s: mp.mpf(1) if s == 0 else mp.atanh(mp.sqrt(s)) / mp.sqrt(s), mp.mpf(0), mp.mpf("0.0405"), 7, scale=2)
