/* CPU emulation (test infrastructure) of the fp64 sum-product check node of the CUDA path (csrc/kernels.cuh bp_check_*):
 * same operations in the same order with fma(); the one difference is the division, which the kernel evaluates with a
 * reciprocal seed + Newton steps (<= 1 ulp) and this file with the correctly rounded '/'.  Included by
 * tests/study_bp_edomain.py into a copy of the C oracle. */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const double ED_EXPC[12] = {0x1.0000000000000p+0, 0x1.0000000000000p+0, 0x1.0000000000011p-1, 0x1.555555555556ap-3, 0x1.555555554f0b8p-5,
                                   0x1.111111110c4b2p-7, 0x1.6c16c188007dap-10, 0x1.a01a01bec709ap-13, 0x1.a01991a047428p-16, 0x1.71ddd953a34f1p-19,
                                   0x1.28b410c11893dp-22, 0x1.af8db1c459b51p-26};
static const double ED_LOGC[8] = {0x1.0000000000000p+1, 0x1.55555555558bap-1, 0x1.99999998be9bbp-2, 0x1.249249cc9eae0p-2,
                                  0x1.c71bf34d80545p-3, 0x1.7476dda759777p-3, 0x1.382eecef5e3a5p-3, 0x1.3bc09a1b49468p-3};
static double ed_max_s = 0;

static inline uint32_t ed_hi(double v) { uint64_t b; memcpy(&b, &v, 8); return (uint32_t)(b >> 32); }
static inline uint32_t ed_lo(double v) { uint64_t b; memcpy(&b, &v, 8); return (uint32_t)b; }
static inline double ed_mk(uint32_t hi, uint32_t lo) { uint64_t b = ((uint64_t)hi << 32) | lo; double v; memcpy(&v, &b, 8); return v; }

static double ed_exp_neg(double z) /* e^-z, z >= 0, clamped at 708 */
{
    const double K0 = -1.4426950408889634, K1 = 6755399441055744.0, K2 = 0.6931471803691238, K3 = 1.9082149292705877e-10;
    const int32_t zh = (int32_t)ed_hi(z);
    double zc = ed_mk((uint32_t)(zh < 0x40862000 ? zh : 0x40862000), ed_lo(z)), t = fma(zc, K0, K1), k = t - K1, r = fma(k, -K2, -zc), p = ED_EXPC[11];
    r = fma(k, -K3, r);
    for (int i = 10; i >= 0; --i) p = fma(p, r, ED_EXPC[i]);
    return ed_mk(ed_hi(p) + (ed_lo(t) << 20), ed_lo(p));
}

/* log(D / N) + shift for 0 < N <= D (both normal numbers); *far: the result is beyond the range the clamp of ed_exp_neg keeps exact */
static double ed_log_ratio(double N, double D, double shift, int *far)
{
    const int32_t hn = (int32_t)ed_hi(N), hd = (int32_t)ed_hi(D);
    const int e0 = (hd - hn + 0x80000) >> 20;
    const double Ns = ed_mk((uint32_t)(hn + e0 * (1 << 20)), ed_lo(N));
    const double w = (D - Ns) / (D + Ns);
    const double s = w * w;
    double p = ED_LOGC[7];
    for (int i = 6; i >= 0; --i) p = fma(p, s, ED_LOGC[i]);
    if (s > ed_max_s) { ed_max_s = s; if (s > 0.0405) { fprintf(stderr, "bp_log_frac: s = %g out of range\n", s); abort(); } }
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
    if (d > 8)
    { /* arbitrary-degree path (tile4.cuh cn4_any): forward / backward values as plain numbers, one division per step */
        int far, anyfar = 0;
        double B = E[d - 1];
        FN[0] = E[0];
        for (int j = 1; j < d - 1; ++j) FN[j] = (FN[j - 1] + E[j]) / fma(FN[j - 1], E[j], 1.0);
        out[d - 1] = ed_log_ratio(FN[d - 2], 1.0, shift, &far); anyfar |= far;
        for (int j = d - 2; j >= 1; --j)
        {
            out[j] = ed_log_ratio(FN[j - 1] + B, fma(FN[j - 1], B, 1.0), shift, &far); anyfar |= far;
            B = (B + E[j]) / fma(B, E[j], 1.0);
        }
        out[0] = ed_log_ratio(B, 1.0, shift, &far); anyfar |= far;
        if (!anyfar)
        {
            for (int j = 0; j < d; ++j) out[j] = copysign(out[j], ((sx ^ ed_hi(x[j])) >> 31) ? -1.0 : 1.0);
            return;
        }
    }
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
