/* CPU emulation (test infrastructure) of the fp64 sum-product check node of the CUDA path (csrc/kernels.cuh bp_check_*):
 * same operations in the same order with fma(); the one difference is the division, which the kernel evaluates with a
 * reciprocal seed + Newton steps (<= 1 ulp) and this file with the correctly rounded '/'.  Included by
 * tests/study_bp_edomain.py into a copy of the C oracle. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const double ED_EXPC[14] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
                                   1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
static const double ED_LOGC[10] = {2.0, 2.0 / 3, 2.0 / 5, 2.0 / 7, 2.0 / 9, 2.0 / 11, 2.0 / 13, 2.0 / 15, 2.0 / 17, 2.0 / 19};

static inline uint32_t ed_hi(double v) { uint64_t b; memcpy(&b, &v, 8); return (uint32_t)(b >> 32); }
static inline uint32_t ed_lo(double v) { uint64_t b; memcpy(&b, &v, 8); return (uint32_t)b; }
static inline double ed_mk(uint32_t hi, uint32_t lo) { uint64_t b = ((uint64_t)hi << 32) | lo; double v; memcpy(&v, &b, 8); return v; }

static double ed_exp_neg(double z) /* e^-z, z >= 0, clamped at 708 */
{
    const double K0 = -1.4426950408889634, K1 = 6755399441055744.0, K2 = 0.6931471803691238, K3 = 1.9082149292705877e-10;
    double zc = fmin(z, 708.0), t = fma(zc, K0, K1), k = t - K1, r = fma(k, -K2, -zc), p = ED_EXPC[13];
    r = fma(k, -K3, r);
    for (int i = 12; i >= 0; --i) p = fma(p, r, ED_EXPC[i]);
    return ed_mk(ed_hi(p) + (ed_lo(t) << 20), ed_lo(p));
}

/* log(D / N) + shift for 0 < N <= D (both normal numbers); *far: the result is beyond the range the clamp of ed_exp_neg keeps exact */
static double ed_log_ratio(double N, double D, double shift, int *far)
{
    const uint32_t hn = ed_hi(N), hd = ed_hi(D);
    int e0 = (int)(hd >> 20) - (int)(hn >> 20);
    float fn, fd;
    uint32_t bn = 0x3f800000u | ((hn & 0xfffffu) << 3), bd = 0x3f800000u | ((hd & 0xfffffu) << 3);
    memcpy(&fn, &bn, 4); memcpy(&fd, &bd, 4);
    e0 += (fd > 1.41421354f * fn) ? 1 : 0;
    e0 -= (fn > 1.41421354f * fd) ? 1 : 0;
    const double Ns = ed_mk(hn + ((uint32_t)e0 << 20), ed_lo(N));
    const double w = (D - Ns) / (D + Ns);
    const double s = w * w;
    double p = ED_LOGC[9];
    for (int i = 8; i >= 0; --i) p = fma(p, s, ED_LOGC[i]);
    *far = e0 > 960;
    return fma(w, p, fma((double)e0, 0.6931471805599453, shift));
}

static double ed_boxplus_exact(double x, double y) /* decoder.h:12-15 */
{
    double m = fmin(fabs(x), fabs(y));
    return copysign(m, (signbit(x) != signbit(y)) ? -1.0 : 1.0) + log((1 + exp(-fabs(x + y))) / (1 + exp(-fabs(x - y))));
}

static void edomain_check(const double *x, int d, double *out)
{
    double E[128], FN[128], FD[128];
    uint32_t sx = 0, mh = 0x7fffffffu;
    if (d == 2) { out[0] = x[1]; out[1] = x[0]; return; }
    for (int j = 0; j < d; ++j) { const uint32_t h = ed_hi(x[j]); sx ^= h; mh = (h & 0x7fffffffu) < mh ? (h & 0x7fffffffu) : mh; }
    const double shift = (mh >= 0x40440000u && mh < 0x7ff00000u) ? ed_mk(mh, 0) - 40.0 : 0.0; /* min |x| (truncated) >= 40: all E <= e^-40 after the shift */
    for (int j = 0; j < d; ++j) E[j] = ed_exp_neg(fabs(x[j]) - shift);
    FN[0] = E[0]; FD[0] = 1.0;
    for (int j = 1; j < d - 1; ++j) { FN[j] = fma(E[j], FD[j - 1], FN[j - 1]); FD[j] = fma(E[j], FN[j - 1], FD[j - 1]); }
    double BN = E[d - 1], BD = 1.0;
    int far, anyfar = 0;
    out[d - 1] = ed_log_ratio(FN[d - 2], FD[d - 2], shift, &far); anyfar |= far;
    for (int j = d - 2; j >= 1; --j)
    {
        const double N = fma(FN[j - 1], BD, BN * FD[j - 1]), D = fma(FN[j - 1], BN, FD[j - 1] * BD);
        out[j] = ed_log_ratio(N, D, shift, &far); anyfar |= far;
        const double bn = fma(E[j], BD, BN), bd = fma(E[j], BN, BD);
        BN = bn; BD = bd;
    }
    out[0] = ed_log_ratio(BN, BD, shift, &far); anyfar |= far;
    if (anyfar || getenv("EDOMAIN_EXACT"))
    {
        /* rare: inputs beyond the clamp decide an output -> the reference's own recursion for this check */
        double F[128], B;
        F[0] = x[0];
        for (int j = 1; j < d; ++j) F[j] = ed_boxplus_exact(F[j - 1], x[j]);
        B = x[d - 1];
        out[d - 1] = F[d - 2];
        for (int j = d - 2; j >= 1; --j) { out[j] = ed_boxplus_exact(F[j - 1], B); B = ed_boxplus_exact(B, x[j]); }
        out[0] = B;
        return;
    }
    for (int j = 0; j < d; ++j) out[j] = copysign(out[j], ((sx ^ ed_hi(x[j])) >> 31) ? -1.0 : 1.0);
}
