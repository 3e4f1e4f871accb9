/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See ldpc_oracle.h for the role of this file.
 * Plain-C restatement of the reference decode path; citations are /root/reference paths. */
#define _GNU_SOURCE
#include "ldpc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* loader — src/core/ldpc.cpp:40-101 (header scan, bit_pos, max degree) and
 *          src/core/sparse.h:91-153 (edge list in file order, rowN/colN in file order)        */
/* ------------------------------------------------------------------------------------------ */

static void parse_int_list(const char *s, int **list, int *n)
{
    /* "while (record >> index) push_back(index)" — ldpc.cpp:60-69 */
    int cap = 16;
    *list = (int *)malloc(sizeof(int) * cap);
    *n = 0;
    for (;;)
    {
        char *end;
        long v = strtol(s, &end, 10);
        if (end == s) break;
        if (*n == cap) { cap *= 2; *list = (int *)realloc(*list, sizeof(int) * cap); }
        (*list)[(*n)++] = (int)v;
        s = end;
    }
}

static int contains(const int *l, int n, int v)
{
    for (int i = 0; i < n; ++i) if (l[i] == v) return 1;
    return 0;
}

static orc_code *load_impl(const char *path, int scan_header, int skip_lines)
{
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    orc_code *c = (orc_code *)calloc(1, sizeof(orc_code));
    char *line = NULL; size_t cap = 0; ssize_t len;
    int ecap = 1024;
    c->e_row = (int *)malloc(sizeof(int) * ecap);
    c->e_col = (int *)malloc(sizeof(int) * ecap);
    c->punct = (int *)malloc(sizeof(int)); c->shorten = (int *)malloc(sizeof(int));
    int in_header = scan_header;
    int maxr = 0, maxc = 0; /* numCols/numRows start at 0, sparse.h:103-104 */
    while ((len = getline(&line, &cap, f)) >= 0)
    {
        if (skip_lines > 0) { --skip_lines; continue; }
        if (in_header)
        {
            char *colon = strchr(line, ':');
            if (colon)
            { /* legacy "key: value" header lines; only puncture/shorten are interpreted, ldpc.cpp:53-72 */
                *colon = 0;
                if (strstr(line, "puncture")) { free(c->punct); parse_int_list(colon + 1, &c->punct, &c->n_punct); }
                else if (strstr(line, "shorten")) { free(c->shorten); parse_int_list(colon + 1, &c->shorten, &c->n_short); }
                continue;
            }
            in_header = 0; /* first line without ':' ends the header, ldpc.cpp:73-76 */
        }
        int r, cc, v;
        int got = sscanf(line, "%d %d %d", &r, &cc, &v);
        if (got < 2) continue; /* blank/garbage line: the reference pushes uninitialised ints (UB); ignored here */
        if (c->nnz == ecap)
        {
            ecap *= 2;
            c->e_row = (int *)realloc(c->e_row, sizeof(int) * ecap);
            c->e_col = (int *)realloc(c->e_col, sizeof(int) * ecap);
        }
        /* every listed entry is a 1 over GF(2) (missing/zero value -> 1, sparse.h:124-128) */
        c->e_row[c->nnz] = r; c->e_col[c->nnz] = cc; c->nnz++;
        if (cc > maxc) maxc = cc;
        if (r > maxr) maxr = r;
    }
    free(line);
    fclose(f);
    c->nc = maxc + 1; c->mc = maxr + 1; /* sparse.h:136-143 */

    /* rowN / colN: per node the edge ids in file order (sparse.h:132-133) */
    c->row_ptr = (int *)calloc(c->mc + 1, sizeof(int));
    c->col_ptr = (int *)calloc(c->nc + 1, sizeof(int));
    for (int e = 0; e < c->nnz; ++e) { c->row_ptr[c->e_row[e] + 1]++; c->col_ptr[c->e_col[e] + 1]++; }
    for (int i = 0; i < c->mc; ++i) c->row_ptr[i + 1] += c->row_ptr[i];
    for (int i = 0; i < c->nc; ++i) c->col_ptr[i + 1] += c->col_ptr[i];
    c->row_edge = (int *)malloc(sizeof(int) * (c->nnz + 1));
    c->col_edge = (int *)malloc(sizeof(int) * (c->nnz + 1));
    int *rfill = (int *)calloc(c->mc, sizeof(int)), *cfill = (int *)calloc(c->nc, sizeof(int));
    for (int e = 0; e < c->nnz; ++e)
    {
        int r = c->e_row[e], cc = c->e_col[e];
        c->row_edge[c->row_ptr[r] + rfill[r]++] = e;
        c->col_edge[c->col_ptr[cc] + cfill[cc]++] = e;
    }
    free(rfill); free(cfill);

    c->max_degree = 0; /* ldpc.cpp:83-87 */
    for (int i = 0; i < c->mc; ++i) { int d = c->row_ptr[i + 1] - c->row_ptr[i]; if (d > c->max_degree) c->max_degree = d; }
    for (int i = 0; i < c->nc; ++i) { int d = c->col_ptr[i + 1] - c->col_ptr[i]; if (d > c->max_degree) c->max_degree = d; }

    /* derived sizes, ldpc.h:47-59; note the list *sizes* are used, duplicates included */
    c->kc = c->nc - c->mc;
    c->nct = c->nc - c->n_punct - c->n_short;
    c->mct = c->mc - c->n_punct;
    c->kct = c->nct - c->mct;
    /* bit_pos: ascending indices neither shortened nor punctured, ldpc.cpp:90-100 */
    c->bit_pos = (int *)malloc(sizeof(int) * (c->nc + 1));
    int nb = 0;
    for (int i = 0; i < c->nc; ++i)
    {
        if (contains(c->shorten, c->n_short, i)) continue;
        if (contains(c->punct, c->n_punct, i)) continue;
        c->bit_pos[nb++] = i;
    }
    /* with duplicate-free in-range lists nb == nct; otherwise the reference would index out of range */
    if (nb < c->nct) c->nct = nb;
    return c;
}

orc_code *orc_load(const char *path) { return load_impl(path, 1, 0); }
orc_code *orc_load_matrix(const char *path, int skip_lines) { return load_impl(path, 0, skip_lines); }

void orc_free(orc_code *c)
{
    if (!c) return;
    free(c->punct); free(c->shorten); free(c->bit_pos); free(c->e_row); free(c->e_col);
    free(c->row_ptr); free(c->row_edge); free(c->col_ptr); free(c->col_edge); free(c);
}

/* ------------------------------------------------------------------------------------------ */
/* decoder — src/decoding/decoder.h:7-20 (sign, jacobian, minsum), decoder.cpp:11-78 (decode) */
/* ------------------------------------------------------------------------------------------ */

static inline int sgn(double x) { return 1 - 2 * (signbit(x) ? 1 : 0); } /* decoder.h:7-10: -0.0 is negative */
static inline double dmin(double a, double b) { return (b < a) ? b : a; }  /* std::min */

static double f_minsum(double x, double y) /* decoder.h:17-20 */
{
    return sgn(x) * sgn(y) * dmin(fabs(x), fabs(y));
}

static double f_jacobian(double x, double y) /* decoder.h:12-15 */
{
    return sgn(x) * sgn(y) * dmin(fabs(x), fabs(y)) + log((1 + exp(-fabs(x + y))) / (1 + exp(-fabs(x - y))));
}

int orc_is_codeword(const orc_code *c, const uint8_t *co) /* decoder.h:47-64 */
{
    for (int i = 0; i < c->mc; ++i)
    {
        uint8_t s = 0;
        for (int k = c->row_ptr[i]; k < c->row_ptr[i + 1]; ++k) s ^= co[c->e_col[c->row_edge[k]]];
        if (s) return 0;
    }
    return 1;
}

int orc_decode(const orc_code *c, const double *llr_in, int iterations, int early_term, int minsum,
               double *llr_out, uint8_t *co)
{
    double (*f)(double, double) = minsum ? f_minsum : f_jacobian;
    double *v2c = (double *)malloc(sizeof(double) * (c->nnz + 1));
    double *c2v = (double *)calloc(c->nnz + 1, sizeof(double));
    double *F = (double *)calloc(c->max_degree + 2, sizeof(double));
    double *B = (double *)calloc(c->max_degree + 2, sizeof(double));
    for (int i = 0; i < c->nc; ++i) { llr_out[i] = 0.0; co[i] = 0; } /* vectors are value-initialised, decoder.h:35-39 */

    for (int e = 0; e < c->nnz; ++e) v2c[e] = llr_in[c->e_col[e]]; /* decoder.cpp:16-19 */

    int I = 0;
    while (I < iterations) /* decoder.cpp:22 */
    {
        for (int i = 0; i < c->mc; ++i) /* CN processing, decoder.cpp:25-45 */
        {
            const int *cn = c->row_edge + c->row_ptr[i];
            int cw = c->row_ptr[i + 1] - c->row_ptr[i];
            if (cw < 2) continue; /* degree 0/1 checks index out of range in the reference (UB); skipped */
            F[0] = v2c[cn[0]];
            B[cw - 1] = v2c[cn[cw - 1]];
            for (int j = 1; j < cw; ++j)
            {
                F[j] = f(F[j - 1], v2c[cn[j]]);
                B[cw - 1 - j] = f(B[cw - j], v2c[cn[cw - j - 1]]);
            }
            c2v[cn[0]] = B[1];
            c2v[cn[cw - 1]] = F[cw - 2];
            for (int j = 1; j < cw - 1; ++j) c2v[cn[j]] = f(F[j - 1], B[j + 1]);
        }
        for (int i = 0; i < c->nc; ++i) /* VN processing and app calc, decoder.cpp:48-64 */
        {
            double out = llr_in[i];
            for (int k = c->col_ptr[i]; k < c->col_ptr[i + 1]; ++k) out += c2v[c->col_edge[k]];
            llr_out[i] = out;
            co[i] = (out <= 0);
            for (int k = c->col_ptr[i]; k < c->col_ptr[i + 1]; ++k) v2c[c->col_edge[k]] = out - c2v[c->col_edge[k]];
        }
        if (early_term && orc_is_codeword(c, co)) break; /* decoder.cpp:66-72: break BEFORE ++I */
        ++I;
    }
    free(v2c); free(c2v); free(F); free(B);
    return I;
}

/* ------------------------------------------------------------------------------------------ */
/* Layered schedule — the legacy tree's layer loop (gpu/device/kernel.cpp:52-75: per layer a check  */
/* update of the layer's checks, then the a-posteriori values, then the early-termination test)    */
/* with the check-node rules of the live decoder (forward/backward recursion, box-plus / min-sum). */
/* This is the SPECIFICATION of the opt-in layered mode of the CUDA path; it differs from the      */
/* flooding reference by design.  Deviations from the legacy kernels, all stated in DESIGN.md:     */
/*   - the posterior of a variable is updated in place by the checks of the layer                  */
/*     (out = (out - c2v_old) + c2v_new) instead of being re-summed over all its edges after every */
/*     layer (kernel.cpp:272-296): mathematically the same value; a layer must not contain two     */
/*     checks sharing a variable (orc_layers_valid)                                                */
/*   - early termination is tested once per iteration, after the last layer (the legacy code tests */
/*     after every layer, kernel.cpp:62-69)                                                        */
/* layers: layer_ptr[nl+1], layer_check[]; the checks of a layer in any order (they are independent).*/
/* ------------------------------------------------------------------------------------------ */
int orc_layers_valid(const orc_code *c, int nl, const int *layer_ptr, const int *layer_check)
{
    int *seen = (int *)malloc(sizeof(int) * c->nc), *used = (int *)calloc(c->mc, sizeof(int));
    int ok = 1;
    for (int i = 0; i < c->nc; ++i) seen[i] = -1;
    for (int l = 0; l < nl && ok; ++l)
        for (int q = layer_ptr[l]; q < layer_ptr[l + 1] && ok; ++q)
        {
            const int chk = layer_check[q];
            if (chk < 0 || chk >= c->mc || used[chk]++) { ok = 0; break; }
            for (int k = c->row_ptr[chk]; k < c->row_ptr[chk + 1]; ++k)
            {
                const int v = c->e_col[c->row_edge[k]];
                if (seen[v] == l) { ok = 0; break; }
                seen[v] = l;
            }
        }
    for (int i = 0; i < c->mc && ok; ++i) if (!used[i]) ok = 0; /* every check in exactly one layer */
    free(seen); free(used);
    return ok;
}

int orc_decode_layered(const orc_code *c, int nl, const int *layer_ptr, const int *layer_check, const double *llr_in, int iterations,
                       int early_term, int minsum, double ms_scale, double *llr_out, uint8_t *co)
{
    double (*f)(double, double) = minsum ? f_minsum : f_jacobian;
    double *c2v = (double *)calloc(c->nnz + 1, sizeof(double)); /* +0: the first visit of a check reads the posterior itself */
    double *F = (double *)calloc(c->max_degree + 2, sizeof(double));
    double *B = (double *)calloc(c->max_degree + 2, sizeof(double));
    double *v = (double *)calloc(c->max_degree + 2, sizeof(double));
    for (int i = 0; i < c->nc; ++i) { llr_out[i] = llr_in[i]; co[i] = 0; }
    int I = 0;
    while (I < iterations)
    {
        for (int l = 0; l < nl; ++l)
            for (int q = layer_ptr[l]; q < layer_ptr[l + 1]; ++q)
            {
                const int *cn = c->row_edge + c->row_ptr[layer_check[q]];
                const int cw = c->row_ptr[layer_check[q] + 1] - c->row_ptr[layer_check[q]];
                if (cw < 2) continue;
                for (int j = 0; j < cw; ++j) v[j] = llr_out[c->e_col[cn[j]]] - c2v[cn[j]]; /* the v2c of kernel.cpp:290-294 */
                F[0] = v[0];
                B[cw - 1] = v[cw - 1];
                for (int j = 1; j < cw; ++j)
                {
                    F[j] = f(F[j - 1], v[j]);
                    B[cw - 1 - j] = f(B[cw - j], v[cw - j - 1]);
                }
                for (int j = 0; j < cw; ++j)
                {
                    double r = (j == 0) ? B[1] : (j == cw - 1) ? F[cw - 2] : f(F[j - 1], B[j + 1]);
                    if (minsum) r *= ms_scale; /* normalised min-sum (1 = plain: exact) */
                    c2v[cn[j]] = r;
                    llr_out[c->e_col[cn[j]]] = v[j] + r;
                }
            }
        for (int i = 0; i < c->nc; ++i) co[i] = (llr_out[i] <= 0); /* kernel.cpp:284 */
        if (early_term && orc_is_codeword(c, co)) break;           /* before ++I, like the live decoder */
        ++I;
    }
    for (int i = 0; i < c->nc; ++i) co[i] = (llr_out[i] <= 0);
    free(c2v); free(F); free(B); free(v);
    return I;
}

/* Layers in the sense of orc_layers_valid by first fit: check i joins the first layer none of whose checks shares a
 * variable with it.  Quasi-cyclic codes come out with (at most) one layer per block row.  Returns the number of layers;
 * layer_of[mc]. */
int orc_auto_layers(const orc_code *c, int *layer_of)
{
    int nl = 0, cap = 16;
    uint8_t **occ = (uint8_t **)malloc(sizeof(uint8_t *) * cap); /* occ[l][v] = variable v already used by layer l */
    for (int i = 0; i < c->mc; ++i)
    {
        int l = 0;
        for (;; ++l)
        {
            if (l == nl)
            {
                if (nl == cap) { cap *= 2; occ = (uint8_t **)realloc(occ, sizeof(uint8_t *) * cap); }
                occ[nl++] = (uint8_t *)calloc(c->nc, 1);
            }
            int clash = 0;
            for (int k = c->row_ptr[i]; k < c->row_ptr[i + 1] && !clash; ++k) clash = occ[l][c->e_col[c->row_edge[k]]];
            if (!clash) break;
        }
        for (int k = c->row_ptr[i]; k < c->row_ptr[i + 1]; ++k) occ[l][c->e_col[c->row_edge[k]]] = 1;
        layer_of[i] = l;
    }
    for (int l = 0; l < nl; ++l) free(occ[l]);
    free(occ);
    return nl;
}

/* BEC decoder — decoder.h:145-155 (vn_update, cn_update), decoder.cpp:91-192 */
static inline uint8_t bec_cn(uint8_t l, uint8_t r)
{
    return (l == ORC_ERASURE || r == ORC_ERASURE) ? ORC_ERASURE : (uint8_t)((l != 0) ^ (r != 0));
}
static inline uint8_t bec_vn(uint8_t l, uint8_t r, uint8_t xi)
{
    return (xi == l || xi == r) ? xi : ORC_ERASURE;
}

int orc_decode_bec(const orc_code *c, const uint8_t *llr_in, const uint8_t *cw_true, int iterations,
                   int early_term, int deg1_compat, uint8_t *llr_out, uint8_t *co)
{
    uint8_t *v2c = (uint8_t *)malloc(c->nnz + 1), *c2v = (uint8_t *)calloc(c->nnz + 1, 1);
    uint8_t *F = (uint8_t *)calloc(c->max_degree + 2, 1), *B = (uint8_t *)calloc(c->max_degree + 2, 1);
    for (int i = 0; i < c->nc; ++i) { llr_out[i] = 0; co[i] = 0; }
    for (int e = 0; e < c->nnz; ++e) v2c[e] = llr_in[c->e_col[e]]; /* decoder.cpp:96-99 */
    int I = 0;
    while (I < iterations)
    {
        for (int i = 0; i < c->mc; ++i) /* decoder.cpp:105-123 */
        {
            const int *cn = c->row_edge + c->row_ptr[i];
            int cw = c->row_ptr[i + 1] - c->row_ptr[i];
            if (cw < 2) continue;
            F[0] = v2c[cn[0]];
            B[cw - 1] = v2c[cn[cw - 1]];
            for (int j = 1; j < cw; ++j)
            {
                F[j] = bec_cn(F[j - 1], v2c[cn[j]]);
                B[cw - 1 - j] = bec_cn(B[cw - j], v2c[cn[cw - j - 1]]);
            }
            c2v[cn[0]] = B[1];
            c2v[cn[cw - 1]] = F[cw - 2];
            for (int j = 1; j < cw - 1; ++j) c2v[cn[j]] = bec_cn(F[j - 1], B[j + 1]);
        }
        for (int i = 0; i < c->nc; ++i) /* decoder.cpp:126-167 */
        {
            const int *vn = c->col_edge + c->col_ptr[i];
            int vw = c->col_ptr[i + 1] - c->col_ptr[i];
            uint8_t xi = cw_true[i];
            if (llr_in[i] != ORC_ERASURE)
            {
                for (int k = 0; k < vw; ++k) v2c[vn[k]] = xi;
                llr_out[i] = xi;
                co[i] = xi;
            }
            else if (vw == 0)
            {
                llr_out[i] = ORC_ERASURE; co[i] = 1; /* isolated erased bit (reference: UB) */
            }
            else if (vw == 1)
            {
                /* reference reads mExMsgF[-1] (decoder.cpp:155-156, UB).  Observed value: 0. */
                v2c[vn[0]] = deg1_compat ? 0 : ORC_ERASURE;
                llr_out[i] = c2v[vn[0]];
                co[i] = (llr_out[i] == ORC_ERASURE) ? 1 : xi;
            }
            else
            {
                F[0] = c2v[vn[0]];
                B[vw - 1] = c2v[vn[vw - 1]];
                for (int j = 1; j < vw; ++j)
                {
                    F[j] = bec_vn(F[j - 1], c2v[vn[j]], xi);
                    B[vw - 1 - j] = bec_vn(B[vw - j], c2v[vn[vw - j - 1]], xi);
                }
                v2c[vn[0]] = B[1];
                v2c[vn[vw - 1]] = F[vw - 2];
                for (int j = 1; j < vw - 1; ++j) v2c[vn[j]] = bec_vn(F[j - 1], B[j + 1], xi);
                llr_out[i] = F[vw - 1];
                /* "-channelInput[i]" is gf2(~v) == 1 for both v (gf2.cpp:5-8) */
                co[i] = (llr_out[i] == ORC_ERASURE) ? 1 : xi;
            }
        }
        if (early_term) /* decoder.cpp:169-186: stop when no output is erased */
        {
            int found = 0;
            for (int i = 0; i < c->nc; ++i) if (llr_out[i] == ORC_ERASURE) { found = 1; break; }
            if (!found) break;
        }
        ++I;
    }
    free(v2c); free(c2v); free(F); free(B);
    return I;
}

/* ------------------------------------------------------------------------------------------ */
/* channel LLR rules — src/sim/channel.cpp                                                    */
/* ------------------------------------------------------------------------------------------ */

void orc_llr_awgn(const orc_code *c, const double *y, double sigma2, double *llr) /* channel.cpp:70-93 */
{
    for (int i = 0; i < c->n_punct; ++i) llr[c->punct[i]] = 0.0;
    for (int i = 0; i < c->n_short; ++i) llr[c->shorten[i]] = 99999.9;
    for (int i = 0; i < c->nct; ++i) llr[c->bit_pos[i]] = 2 * y[i] / sigma2;
}

void orc_llr_bsc(const orc_code *c, const uint8_t *y, double eps, double *llr) /* channel.cpp:137-162 */
{
    const double delta = log((1 - eps) / eps);
    for (int i = 0; i < c->n_punct; ++i) llr[c->punct[i]] = 0.0;
    for (int i = 0; i < c->n_short; ++i) llr[c->shorten[i]] = delta;
    for (int i = 0; i < c->nct; ++i) llr[c->bit_pos[i]] = delta * (1 - 2 * (int)y[i]);
}

void orc_llr_bec(const orc_code *c, const uint8_t *y, const uint8_t *cw, uint8_t *llr) /* channel.cpp:207-229 */
{
    for (int i = 0; i < c->n_punct; ++i) llr[c->punct[i]] = ORC_ERASURE;
    /* reference indexes mX (length nct) with the code index s (latent bug, channel.cpp:221); the true bit is meant */
    for (int i = 0; i < c->n_short; ++i) llr[c->shorten[i]] = cw[c->shorten[i]];
    for (int i = 0; i < c->nct; ++i) llr[c->bit_pos[i]] = y[i];
}

int orc_count_bit_errors(const orc_code *c, const uint8_t *co, const uint8_t *cw) /* ldpcsim.cpp:184-188 */
{
    int n = 0;
    for (int i = 0; i < c->nct; ++i) n += (co[c->bit_pos[i]] != cw[c->bit_pos[i]]);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* GF(2) helpers — src/core/sparse.h                                                          */
/* ------------------------------------------------------------------------------------------ */

void orc_multiply_left(const orc_code *g, const uint8_t *left, uint8_t *result) /* sparse.h:162-172: result += left*G */
{
    for (int j = 0; j < g->nc; ++j)
        for (int k = g->col_ptr[j]; k < g->col_ptr[j + 1]; ++k) result[j] ^= (left[g->e_row[g->col_edge[k]]] & 1);
}

void orc_multiply_right(const orc_code *h, const uint8_t *right, uint8_t *result) /* sparse.h:196-206 */
{
    for (int i = 0; i < h->mc; ++i)
        for (int k = h->row_ptr[i]; k < h->row_ptr[i + 1]; ++k) result[i] ^= (right[h->e_col[h->row_edge[k]]] & 1);
}

int orc_rank(const orc_code *h) /* value of sparse.h:227-294 (rank is algorithm independent): dense bit-packed elimination */
{
    int words = (h->nc + 63) / 64;
    uint64_t *m = (uint64_t *)calloc((size_t)h->mc * words, sizeof(uint64_t));
    for (int e = 0; e < h->nnz; ++e) m[(size_t)h->e_row[e] * words + h->e_col[e] / 64] ^= 1ull << (h->e_col[e] % 64);
    int rank = 0;
    for (int col = 0; col < h->nc && rank < h->mc; ++col)
    {
        int w = col / 64; uint64_t bit = 1ull << (col % 64);
        int piv = -1;
        for (int r = rank; r < h->mc; ++r) if (m[(size_t)r * words + w] & bit) { piv = r; break; }
        if (piv < 0) continue;
        if (piv != rank)
            for (int k = 0; k < words; ++k) { uint64_t t = m[(size_t)piv * words + k]; m[(size_t)piv * words + k] = m[(size_t)rank * words + k]; m[(size_t)rank * words + k] = t; }
        for (int r = rank + 1; r < h->mc; ++r)
            if (m[(size_t)r * words + w] & bit)
                for (int k = w; k < words; ++k) m[(size_t)r * words + k] ^= m[(size_t)rank * words + k];
        ++rank;
    }
    free(m);
    return rank;
}

/* ------------------------------------------------------------------------------------------ */
/* specification of the counter-based channel (new; replaces mt19937_64 of channel.cpp:10-11)  */
/* ------------------------------------------------------------------------------------------ */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r)
    {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void philox_block(uint64_t seed, uint32_t point, uint32_t stream, uint64_t frame, uint32_t j, uint32_t out[4])
{
    uint32_t ctr[4] = {j, (uint32_t)frame, (uint32_t)(frame >> 32), (point & 0xFFFFFFu) | (stream << 24)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}

/* Standard-normal pair of the channel specification, version 2: Box-Muller evaluated in IEEE binary32 with nothing but
 * correctly rounded operations (+, *, fma, 1/x, sqrt) in a fixed order, so that this C restatement and the CUDA kernel
 * (libldpc_b200/csrc/kernels.cuh normal_pair) produce the SAME bits.  Input: two Philox words.
 *   radius: 40 uniform bits U = wr:wa[7:0], u1 = (U + 1/2) 2^-40 in (0,1) (tails out to 7.4 sigma);
 *           -ln u1 = e ln 2 - ln f with f in (0.7071, 1.4142] the truncated 24-bit mantissa of 2U+1,
 *           ln f = 2 atanh(s), s = (f-1)/(f+1), odd series up to s^9 (|s| <= 0.1716: truncation 2e-9 relative)
 *   angle : 24 bits: octant wa[31:29], phi = (wa[28:8] + 1/2) 2^-21 pi/4 in (0, pi/4), Taylor sine / cosine (<= 2e-9),
 *           octants by swap / sign (the grid of phi is symmetric, so the angle stays exactly uniform on it). */
static void normal_pair_v2(uint32_t wr, uint32_t wa, float z[2])
{
    const uint64_t T = ((((uint64_t)wr << 8) | (uint64_t)(wa & 0xFFu)) << 1) | 1ull; /* 2U+1 < 2^41 */
    const int lz = __builtin_clzll(T) - 23;                                           /* 0..40 */
    const uint32_t m = (uint32_t)((T << lz) >> 17);                                   /* top 24 bits, truncated */
    float f = (float)m * 0x1p-23f;                                                    /* [1,2), exact */
    int e = lz + 1;                                                                   /* u1 ~ f 2^-e */
    if (f > 1.41421354f) { f = f * 0.5f; e -= 1; }
    const float g = f - 1.0f;                                                         /* exact */
    const float s = g * (1.0f / (2.0f + g));
    const float s2 = s * s;
    float q = 0x1.c71c72p-4f;                                                         /* 1/9 */
    q = fmaf(q, s2, 0x1.24924ap-3f);                                                  /* 1/7 */
    q = fmaf(q, s2, 0x1.99999ap-3f);                                                  /* 1/5 */
    q = fmaf(q, s2, 0x1.555556p-2f);                                                  /* 1/3 */
    q = fmaf(q, s2, 1.0f);
    const float lnf = (2.0f * s) * q;
    const float n = fmaf((float)e, 0x1.62e430p-1f, -lnf);                             /* -ln u1 > 0 */
    const float r = sqrtf(2.0f * n);
    const uint32_t a = wa >> 8, oct = a >> 21;
    const float phi = ((float)(a & 0x1FFFFFu) + 0.5f) * 0x1.921fb6p-22f;              /* pi/4 * 2^-21 */
    const float x2 = phi * phi;
    float sp = 0x1.71de3ap-19f;                                                       /* 1/9! */
    sp = fmaf(sp, x2, -0x1.a01a02p-13f);                                              /* -1/7! */
    sp = fmaf(sp, x2, 0x1.111112p-7f);                                                /* 1/5! */
    sp = fmaf(sp, x2, -0x1.555556p-3f);                                               /* -1/3! */
    const float sn0 = fmaf(phi * x2, sp, phi);
    float cp = -0x1.27e4fcp-22f;                                                      /* -1/10! */
    cp = fmaf(cp, x2, 0x1.a01a02p-16f);                                               /* 1/8! */
    cp = fmaf(cp, x2, -0x1.6c16c2p-10f);                                              /* -1/6! */
    cp = fmaf(cp, x2, 0x1.555556p-5f);                                                /* 1/4! */
    cp = fmaf(cp, x2, -0.5f);
    const float cs0 = fmaf(cp, x2, 1.0f);
    float c = (oct & 1u) ? sn0 : cs0, sn = (oct & 1u) ? cs0 : sn0;                    /* theta = pi/2 - phi in odd octants */
    if (oct & 2u) { const float t = c; c = -sn; sn = t; }                             /* + pi/2 */
    if (oct & 4u) { c = -c; sn = -sn; }                                               /* + pi */
    z[0] = r * c;
    z[1] = r * sn;
}

void orc_normal_from_words(uint32_t wr, uint32_t wa, double z[2])
{
    float a[2];
    normal_pair_v2(wr, wa, a);
    z[0] = a[0]; z[1] = a[1];
}

/* the four normals of Philox block j of a frame (transmitted indices j, j+nb, j+2nb, j+3nb with nb = ceil(nct/4) blocks per
 * frame: a warp of the kernel then touches neighbouring positions): words (0,1) -> first pair, (2,3) -> second */
void orc_normal_block(uint64_t seed, uint32_t point, uint64_t frame, uint32_t j, double z[4])
{
    uint32_t x[4];
    float a[2], b[2];
    philox_block(seed, point, 0, frame, j, x);
    normal_pair_v2(x[0], x[1], a);
    normal_pair_v2(x[2], x[3], b);
    z[0] = a[0]; z[1] = a[1]; z[2] = b[0]; z[3] = b[1];
}

static uint32_t prob_threshold(double p)
{
    double t = floor(p * 4294967296.0);
    if (t <= 0) return 0;
    if (t >= 4294967295.0) return 4294967295u;
    return (uint32_t)t;
}

void orc_channel_frame(const orc_code *c, const orc_code *g, int kind, double x, uint64_t seed,
                       uint32_t point, uint64_t frame, uint8_t *cw, double *llr_f64, uint8_t *llr_u8)
{
    memset(cw, 0, c->nc);
    if (g)
    { /* fresh random information word per frame (stream 1), cw = u*G (channel.cpp:44-52 without the
         accumulate-into-previous-codeword quirk of sparse.h:169) */
        uint8_t *u = (uint8_t *)malloc(g->mc);
        for (int k = 0; k < g->mc; ++k)
        {
            uint32_t w[4];
            philox_block(seed, point, 1, frame, (uint32_t)k >> 7, w);
            u[k] = (w[(k >> 5) & 3] >> (k & 31)) & 1;
        }
        orc_multiply_left(g, u, cw);
        free(u);
    }
    if (kind == ORC_AWGN)
    {
        double sigma2 = pow(10, -x / 10), sigma = sqrt(sigma2); /* channel.cpp:39-41 */
        double *y = (double *)malloc(sizeof(double) * (c->nct + 1));
        const int nb = (c->nct + 3) / 4; /* Philox block q serves the transmitted indices q, q+nb, q+2nb, q+3nb */
        for (int q = 0; q < nb; ++q)
        {
            double z[4];
            orc_normal_block(seed, point, frame, (uint32_t)q, z);
            for (int k = 0; k < 4; ++k)
            {
                const int t = q + k * nb;
                if (t >= c->nct) continue;
                double xs = 1 - 2 * (int)cw[c->bit_pos[t]]; /* channel.cpp:58 */
                y[t] = z[k] * sigma + xs;                    /* channel.cpp:66: normal(0,sigma)() + x */
            }
        }
        /* LLR rule of channel.cpp:70-93 with 2y/sigma^2 evaluated as y * (2/sigma^2), the factor rounded once (what the kernel does) */
        {
            const double scale = 2.0 / sigma2;
            for (int i = 0; i < c->n_punct; ++i) llr_f64[c->punct[i]] = 0.0;
            for (int i = 0; i < c->n_short; ++i) llr_f64[c->shorten[i]] = 99999.9;
            for (int t = 0; t < c->nct; ++t) llr_f64[c->bit_pos[t]] = y[t] * scale;
        }
        free(y);
    }
    else
    {
        uint32_t thr = prob_threshold(x);
        uint8_t *y = (uint8_t *)malloc(c->nct + 1);
        const int nb = (c->nct + 3) / 4; /* value k of Philox block q belongs to the transmitted index q + k*nb */
        for (int t = 0; t < c->nct; ++t)
        {
            uint32_t w[4];
            philox_block(seed, point, 0, frame, (uint32_t)(t % nb), w);
            int hit = w[t / nb] < thr;
            uint8_t b = cw[c->bit_pos[t]];
            if (kind == ORC_BSC) y[t] = b ^ (uint8_t)hit;            /* channel.cpp:131 */
            else y[t] = hit ? ORC_ERASURE : b;                       /* channel.cpp:201 */
        }
        if (kind == ORC_BSC) orc_llr_bsc(c, y, x, llr_f64);
        else orc_llr_bec(c, y, cw, llr_u8);
        free(y);
    }
}

/* M-ASK with bit-metric decoding — the legacy tree's higher-order modulation (constellation: gpu/sim/ldpcsim.cpp:6-20; bit
 * mapper / labels: :64-140; encode_all0 + map_c_to_x: :240-271; calc_llrs: :273-323 = gpu/device/kernel.cpp:141-219).
 * Specification of the counter-based version: scrambling bits c[v] = bit v of Philox stream 2 (word v/32 of block v/128), symbol
 * i carries the bits c[bm[k][i]] (level k = most significant first), y = X[x] + sigma*z with z = value i / nb of Philox block
 * i % nb of stream 0 (nb = ceil(n_sym/4)), bit-metric LLR log(sum_{bit=0} / sum_{bit=1}) of exp(-(y-X_j)^2/(2 sigma^2)) pX_j,
 * +-inf clipped to +-9999.9, multiplied by (1 - 2c) so that the decoder sees the all-zero codeword; punctured positions 0,
 * shortened 99999.9 (conventions of the live channel, src/sim/channel.cpp:77,84).  cw_out receives the scrambling bits. */
void orc_channel_frame_ask(const orc_code *c, int M, const int *labels, const int *bm, double snr, uint64_t seed, uint32_t point,
                           uint64_t frame, uint8_t *cw_out, double *llr)
{
    int bits = 0;
    while ((1 << bits) < M) ++bits;
    const int n_sym = c->nct / bits, nb = (n_sym + 3) / 4;
    const double sigma2 = pow(10, -snr / 10), sigma = sqrt(sigma2);
    double *X = (double *)malloc(sizeof(double) * M);
    int *rev = (int *)malloc(sizeof(int) * M);
    double m = 0;
    for (int j = 0; j < M; ++j) { X[j] = (double)-M + 1 + 2 * j; m += X[j] * X[j] * (1.0 / M); } /* ldpcsim.cpp:10-14 */
    for (int j = 0; j < M; ++j) { X[j] = X[j] / sqrt(m); rev[labels[j]] = j; }                    /* :16-19, :80-84 */
    memset(cw_out, 0, c->nc);
    for (int i = 0; i < c->n_punct; ++i) llr[c->punct[i]] = 0.0;
    for (int i = 0; i < c->n_short; ++i) llr[c->shorten[i]] = 99999.9;
    for (int i = 0; i < n_sym; ++i)
    {
        int tmp = 0;
        for (int k = 0; k < bits; ++k)
        {
            const int v = bm[k * n_sym + i];
            uint32_t w[4];
            philox_block(seed, point, 2, frame, (uint32_t)v >> 7, w);
            cw_out[v] = (w[(v >> 5) & 3] >> (v & 31)) & 1;
            tmp += cw_out[v] << (bits - 1 - k); /* map_c_to_x, ldpcsim.cpp:258-270 */
        }
        double z[4];
        orc_normal_block(seed, point, frame, (uint32_t)(i % nb), z);
        const double y = z[i / nb] * sigma + X[rev[tmp]];
        for (int k = 0; k < bits; ++k)
        {
            double t0 = 0, t1 = 0;
            for (int j = 0; j < M; ++j)
            {
                const double e = exp(-(y - X[j]) * (y - X[j]) / (2 * sigma2)) * (1.0 / M);
                if (labels[j] & (1 << (bits - 1 - k))) t1 += e; else t0 += e;
            }
            double val = log(t0 / t1);
            if (isinf(val)) val = val > 0 ? 9999.9 : -9999.9; /* ldpcsim.h:59-60 */
            const int v = bm[k * n_sym + i];
            llr[v] = val * (1 - 2 * (int)cw_out[v]);
        }
    }
    free(X); free(rev);
}

void orc_sim_point(const orc_code *c, const orc_code *g, int kind, int minsum, int iterations, int early_term,
                   int bec_deg1_compat, double x, uint64_t seed, uint32_t point, uint64_t frame0,
                   uint64_t nframes, int threads, uint64_t counters[4])
{
    uint64_t fec = 0, bec = 0, iters = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads) reduction(+ : fec, bec, iters)
    {
        uint8_t *cw = (uint8_t *)malloc(c->nc), *co = (uint8_t *)malloc(c->nc);
        uint8_t *lu = (uint8_t *)malloc(c->nc), *lou = (uint8_t *)malloc(c->nc);
        double *lf = (double *)malloc(sizeof(double) * c->nc), *lo = (double *)malloc(sizeof(double) * c->nc);
#pragma omp for schedule(dynamic, 4)
        for (uint64_t t = 0; t < nframes; ++t)
        { /* frame loop body, ldpcsim.cpp:160-190 */
            orc_channel_frame(c, g, kind, x, seed, point, frame0 + t, cw, lf, lu);
            int it = (kind == ORC_BEC) ? orc_decode_bec(c, lu, cw, iterations, early_term, bec_deg1_compat, lou, co)
                                       : orc_decode(c, lf, iterations, early_term, minsum, lo, co);
            iters += (uint64_t)it;
            int be = orc_count_bit_errors(c, co, cw);
            if (be > 0) { bec += (uint64_t)be; ++fec; }
        }
        free(cw); free(co); free(lu); free(lou); free(lf); free(lo);
    }
    counters[0] = fec; counters[1] = bec; counters[2] = nframes; counters[3] = iters;
}
