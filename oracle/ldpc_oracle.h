/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the heat1q/libldpc decode hot path (loader -> channel LLR rules ->
 * flooding BP / min-sum / BEC decode -> error accounting).  Every function cites the reference
 * file:line it follows (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (libldpc_b200/libldpc.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path (its own test binary
 * only covers GF(2) ops / rank / G*H^T, tests/ldpctest.cpp:5-76), so this restatement is pinned
 * against outputs of the unmodified reference itself, compiled here by oracle/Makefile into
 * oracle/_ref/ (dump_ref, libldpc_ref.so); the resulting vectors are committed under tests/golden/
 * together with the generating script tests/golden/make_golden.py.
 *
 * The Philox channel (orc_philox4x32_10 / orc_channel_frame) is NOT reference behaviour (the
 * reference uses mt19937_64, src/sim/channel.cpp:10-11,33); it is the executable specification of
 * the new counter-based channel so that the CUDA channel kernel can be checked on the CPU.
 */
#ifndef LDPC_ORACLE_H
#define LDPC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ERASURE 69 /* 'E', src/core/functions.h:105 */

typedef struct orc_code
{
    int nc, mc, nnz;          /* src/core/ldpc.h:47-53 */
    int n_punct, n_short;
    int *punct, *shorten;     /* header lists, src/core/ldpc.cpp:49-76 */
    int nct, mct, kct, kc;    /* src/core/ldpc.h:49,55-59 */
    int max_degree;           /* src/core/ldpc.cpp:83-87 */
    int *bit_pos;             /* [nct] src/core/ldpc.cpp:90-100 */
    int *e_row, *e_col;       /* [nnz] edges in file order, src/core/sparse.h:115-140 */
    int *row_ptr, *row_edge;  /* rowN: edge ids per row, file order, src/core/sparse.h:133,151-152 */
    int *col_ptr, *col_edge;  /* colN: edge ids per col, file order, src/core/sparse.h:132,149-150 */
} orc_code;

/* loader: returns NULL on failure (reference prints and exit()s, src/core/ldpc.cpp:16-20). */
orc_code *orc_load(const char *path);
/* generator-matrix style file: no header skipping (src/core/ldpc.cpp:103-106). */
orc_code *orc_load_matrix(const char *path, int skip_lines);
void orc_free(orc_code *c);

/* flooding BP (minsum=0: pairwise Jacobian box-plus) / min-sum decoder, src/decoding/decoder.cpp:11-78.
 * llr_in/llr_out are full length nc.  Returns the reference's 0-based iteration count. */
int orc_decode(const orc_code *c, const double *llr_in, int iterations, int early_term, int minsum,
               double *llr_out, uint8_t *co);
/* syndrome test on hard decisions, src/decoding/decoder.h:47-64 */
int orc_is_codeword(const orc_code *c, const uint8_t *co);
/* BEC decoder, src/decoding/decoder.cpp:91-192.  deg1_compat=1 reproduces the observed outcome of the
 * reference's out-of-bounds read for erased degree-1 variable nodes (message 0); 0 sends an erasure. */
int orc_decode_bec(const orc_code *c, const uint8_t *llr_in, const uint8_t *cw, int iterations,
                   int early_term, int deg1_compat, uint8_t *llr_out, uint8_t *co);

/* channel LLR rules (noise supplied by the caller), src/sim/channel.cpp:70-93,137-162,207-229 */
void orc_llr_awgn(const orc_code *c, const double *y /*[nct]*/, double sigma2, double *llr /*[nc]*/);
void orc_llr_bsc(const orc_code *c, const uint8_t *y /*[nct]*/, double eps, double *llr /*[nc]*/);
void orc_llr_bec(const orc_code *c, const uint8_t *y /*[nct]*/, const uint8_t *cw /*[nc]*/, uint8_t *llr /*[nc]*/);

/* bit errors over transmitted positions, src/sim/ldpcsim.cpp:184-188 */
int orc_count_bit_errors(const orc_code *c, const uint8_t *co, const uint8_t *cw);

/* GF(2) helpers of the C ABI: src/core/sparse.h:162-187 (encode), :196-218 (syndrome), :227-294 (rank) */
void orc_multiply_left(const orc_code *g, const uint8_t *left, uint8_t *result /* accumulated into */);
void orc_multiply_right(const orc_code *h, const uint8_t *right, uint8_t *result /* accumulated into */);
int orc_rank(const orc_code *h);

/* ---- specification of the new counter-based channel (not reference behaviour) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* channel kinds */
enum { ORC_AWGN = 1, ORC_BSC = 2, ORC_BEC = 3 };
/* Generates one frame: codeword (all-zero unless g != NULL), then the decoder input.
 * AWGN/BSC write llr_f64[nc]; BEC writes llr_u8[nc].  cw[nc] always written. */
void orc_channel_frame(const orc_code *c, const orc_code *g, int kind, double x, uint64_t seed,
                       uint32_t point, uint64_t frame, uint8_t *cw, double *llr_f64, uint8_t *llr_u8);
/* the four standard normals of Philox block j of (seed, point, frame) — the AWGN channel uses them for the transmitted
 * indices j, j+nb, j+2nb, j+3nb, nb = ceil(nct/4) — Box-Muller in
 * exactly specified binary32 arithmetic (the CUDA kernel produces the same bits) */
void orc_normal_block(uint64_t seed, uint32_t point, uint64_t frame, uint32_t j, double z[4]);
/* the same from two raw 32-bit words (radius word, angle word) — for accuracy tests of the evaluation */
void orc_normal_from_words(uint32_t wr, uint32_t wa, double z[2]);

/* Monte-Carlo point over frames [frame0, frame0+nframes): counters = {fec, bec, frames, iters}
 * following src/sim/ldpcsim.cpp:160-190 (no stop rule: every frame is counted).  threads>1 uses OpenMP. */
void orc_sim_point(const orc_code *c, const orc_code *g, int kind, int minsum, int iterations, int early_term,
                   int bec_deg1_compat, double x, uint64_t seed, uint32_t point, uint64_t frame0,
                   uint64_t nframes, int threads, uint64_t counters[4]);

#ifdef __cplusplus
}
#endif
/* layered schedule (specification of the opt-in layered mode; legacy tree gpu/device/kernel.cpp:52-75) */
int orc_layers_valid(const orc_code *c, int nl, const int *layer_ptr, const int *layer_check);
int orc_decode_layered(const orc_code *c, int nl, const int *layer_ptr, const int *layer_check, const double *llr_in, int iterations,
                       int early_term, int minsum, double ms_scale, double *llr_out, uint8_t *co);
int orc_auto_layers(const orc_code *c, int *layer_of);

/* M-ASK with bit-metric decoding (legacy tree gpu/sim/ldpcsim.cpp:6-20,240-323): labels[M], bm[log2M][nct/log2M] variable ids */
void orc_channel_frame_ask(const orc_code *c, int M, const int *labels, const int *bm, double snr, uint64_t seed, uint32_t point,
                           uint64_t frame, uint8_t *cw_out, double *llr);

#endif
