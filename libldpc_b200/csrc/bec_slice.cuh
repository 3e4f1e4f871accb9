// Bit-sliced erasure-channel simulation kernel: BEC channel -> erasure message passing -> error accounting for the
// sweep path (reference: channel_bec, src/sim/channel.cpp:177-229; ldpc_decoder_bec::decode, src/decoding/decoder.cpp:91-192;
// vn_update / cn_update, src/decoding/decoder.h:145-155; accounting, src/sim/ldpcsim.cpp:184-190).
//
// A message of the reference is 'E' or a bit.  Relative to the TRUE bit of the variable an edge belongs to, a message is
// therefore: unknown, known-and-right, or known-and-wrong.  Two bit-planes per edge describe it, "known" K and "wrong" W
// (W subset of K), 32 frames per 32-bit word (frame f = bit f), and the reference's forward/backward recursions collapse to
// word-wide boolean folds that give the same result bit for bit (the codeword satisfies every check, so the XOR of the other
// inputs of a check is the true bit of the target XOR the parity of their wrongness):
//   check            : K'_j = all OTHER inputs known;  W'_j = K'_j & XOR of the other inputs' W          (cn_update)
//   received variable: sends its true bit on every edge: K' = 1, W' = 0
//   erased variable, degree >= 3: K'_j = some OTHER input is known-and-right, W'_j = 0                   (vn_update: "the
//            true bit if any input equals it"); posterior known iff some input is known-and-right
//   erased, degree 2 : the two raw inputs are swapped (K, W travel along); posterior as above
//   erased, degree 1 : sends 0 (the reference's out-of-bounds read): K' = 1, W' = true bit; 'E' with
//            bec_deg1_compat = 0; posterior known iff the input is known (right or wrong: the test is `== 'E'`)
// Wrong bits only ever originate at erased degree-1 variables whose true bit is 1, i.e. with a generator matrix (-G) and the
// compatibility rule; otherwise W is identically zero and the plane is not kept at all (WP = false).  The decision is the true
// bit unless the posterior is 'E', then 1 ("wrong bit", decoder.cpp:165, gf2.cpp:5-8): a bit error iff erased and true bit 0.
// Early termination (no 'E' left among the posteriors, decoder.cpp:169-186) is per frame: knowledge only grows, so a
// finished frame is a fixed point and simply rides along until its 32-frame word retires; the reference's return value
// (iterations completed BEFORE the successful one) is the number of iterations after which the frame still had erasures.
//
// One group of TPG threads owns one 32-frame word: messages live in shared memory in place (a check / a variable reads
// and rewrites only its own edges), groups synchronise on their own named barrier and pull words independently.
// Message slots are numbered by a proper 32-colouring of the edges (BecSliceLayout, code.cpp): slot % 32 is the colour, and
// the edges a warp touches in one access — edge k of 32 consecutive checks, or of 32 consecutive variables — all have
// different colours, so every shared-memory access of both phases is free of bank conflicts for ANY code.
#pragma once
#include "kernels.cuh"

namespace b200
{
    struct BecSliceParams
    {
        const int32_t *row_ptr, *col_ptr;
        const uint16_t *row_slot, *col_slot; // check -> slots of its edges / variable -> slots of its edges (file order)
        const int32_t *tx_var, *punct, *shorten;
        const uint8_t *tx_flag; // [nc] 1 = transmitted position (counts towards bit / frame errors)
        int nc, mc, nnz, nct, n_punct, n_short, n_slots;
        int max_iter, early_term, deg1_compat;
        uint32_t thr;
        uint64_t seed;
        uint32_t point;
        uint64_t frame0, n_frames;
        unsigned long long *counters; // [5] fec, bec, frames, sum(ret iters), sum(executed iterations)
        int groups_per_cta, group_words;
        // generator matrix by column (-G: random codewords u*G, src/sim/channel.cpp:177-191); g_rows = 0 -> all-zero word
        const int32_t *g_col_ptr, *g_row;
        int g_rows, g_cols;
        // per-error diagnostics log (may be null)
        unsigned long long *err_log, *err_count;
        unsigned long long err_cap;
    };

    constexpr int BEC_TPG = 128; // threads per 32-frame word

    __device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(BEC_TPG) : "memory"); }

    // GEN: generator matrix and / or error log in use (true-bit plane, optional wrong plane, per-frame bookkeeping);
    // GEN = false is the lean all-zero-codeword sweep.
    template <bool GEN>
    __global__ void __launch_bounds__(1024, 1) bec_slice_kernel(const BecSliceParams p)
    {
        extern __shared__ __align__(16) uint32_t bs_smem[];
        __shared__ unsigned long long s_cnt[5];
        const int tid = threadIdx.x, gs = tid / BEC_TPG, gt = tid % BEC_TPG, lane = tid & 31;
        const bool has_g = GEN && p.g_rows > 0, logging = GEN && p.err_log != nullptr;
        const bool wp = has_g && p.deg1_compat; // wrong plane kept
        // per group: K[n_slots] | er[nc] | red[8] | (GEN:) xi[nc] | W[n_slots] | pe[nc] | itc[32]
        uint32_t *K = bs_smem + (size_t)gs * p.group_words, *er = K + p.n_slots, *red = er + p.nc;
        uint32_t *xi = red + 8, *W = xi + p.nc, *pe = W + p.n_slots, *itc = pe + p.nc;
        if (tid < 5) s_cnt[tid] = 0;
        __syncthreads();

        const uint64_t n_words = (p.n_frames + 31) / 32;
        for (uint64_t w = (uint64_t)blockIdx.x * p.groups_per_cta + gs; w < n_words; w += (uint64_t)gridDim.x * p.groups_per_cta)
        {
            const uint64_t f0 = w * 32;
            const uint32_t valid = (p.n_frames - f0 >= 32) ? 0xFFFFFFFFu : ((1u << (uint32_t)(p.n_frames - f0)) - 1u);

            if constexpr (GEN)
            {
                if (has_g)
                { // true-bit plane xi[v]: bit f = codeword bit of variable v in frame f0+f, cw = u*G with the information word
                  // of Philox stream 1 (same words as the tile kernel and the channel kernel).  Scratch: the K region.
                    const int nblk = (p.g_rows + 127) >> 7, uw = 4 * nblk;
                    uint32_t *tmp = K, *U = K + 32 * uw;
                    for (int i = gt; i < 32 * nblk; i += BEC_TPG)
                    {
                        const int f = i & 31, blk = i >> 5;
                        const u32x4 r = channel_block(p.seed, p.point, 1, p.frame0 + f0 + f, (uint32_t)blk);
                        tmp[f * uw + 4 * blk] = r.x; tmp[f * uw + 4 * blk + 1] = r.y; tmp[f * uw + 4 * blk + 2] = r.z; tmp[f * uw + 4 * blk + 3] = r.w;
                    }
                    group_barrier(gs + 1);
                    for (int r = gt; r < p.g_rows; r += BEC_TPG)
                    {
                        uint32_t word = 0;
#pragma unroll 8
                        for (int f = 0; f < 32; ++f) word |= ((tmp[f * uw + (r >> 5)] >> (r & 31)) & 1u) << f;
                        U[r] = word;
                    }
                    group_barrier(gs + 1);
                    for (int v = gt; v < p.nc; v += BEC_TPG)
                    {
                        uint32_t x = 0;
                        if (v < p.g_cols)
                            for (int q = p.g_col_ptr[v]; q < p.g_col_ptr[v + 1]; ++q) x ^= U[p.g_row[q]];
                        xi[v] = x;
                    }
                    group_barrier(gs + 1);
                }
                if (logging && gt < 32) itc[gt] = 0;
            }

            // ---- channel: er[v] bit f = frame f0+f sees an erasure at variable v (channel.cpp:193-229) ----
            for (int i = gt; i < p.n_punct; i += BEC_TPG) er[p.punct[i]] = 0xFFFFFFFFu; // punctured: 'E'
            for (int i = gt; i < p.n_short; i += BEC_TPG) er[p.shorten[i]] = 0u;         // shortened: the true bit
            if (gt < 8) red[gt] = 0;
            const int nblk = (p.nct + 3) >> 2;
            for (int q = gt; q < nblk; q += BEC_TPG)
            {
                uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
#pragma unroll 4
                for (int f = 0; f < 32; ++f)
                { // same stream as the byte-wise path: counter (block q, frame)
                    const u32x4 r = channel_block(p.seed, p.point, 0, p.frame0 + f0 + f, (uint32_t)q);
                    e0 |= (r.x < p.thr ? 1u : 0u) << f;
                    e1 |= (r.y < p.thr ? 1u : 0u) << f;
                    e2 |= (r.z < p.thr ? 1u : 0u) << f;
                    e3 |= (r.w < p.thr ? 1u : 0u) << f;
                }
                // value k of block q <-> transmitted index q + k*nblk (channel specification): neighbouring threads, neighbouring words
                er[p.tx_var[q]] = e0;
                if (q + nblk < p.nct) er[p.tx_var[q + nblk]] = e1;
                if (q + 2 * nblk < p.nct) er[p.tx_var[q + 2 * nblk]] = e2;
                if (q + 3 * nblk < p.nct) er[p.tx_var[q + 3 * nblk]] = e3;
            }
            group_barrier(gs + 1);
            // ---- v2c of iteration 0 = the channel value (decoder.cpp:96-99): known iff received, never wrong ----
            for (int v = gt; v < p.nc; v += BEC_TPG)
            {
                const uint32_t k = ~er[v];
                for (int q = p.col_ptr[v]; q < p.col_ptr[v + 1]; ++q)
                {
                    const int s = p.col_slot[q];
                    K[s] = k;
                    if (GEN && wp) W[s] = 0u;
                }
            }
            group_barrier(gs + 1);

            uint32_t ret_sum = 0, err_bits = 0, err_or = 0, still = valid;
            int it = 0;
            for (; it < p.max_iter; ++it)
            {
                // ---- check nodes (decoder.cpp:105-123) ----
                for (int c = gt; c < p.mc; c += BEC_TPG)
                {
                    const int b = p.row_ptr[c], e = p.row_ptr[c + 1];
                    if (e - b < 2) continue;
                    uint32_t one = 0, two = 0, tw = 0; // frames with >= 1 / >= 2 unknown inputs; parity of the wrong inputs
                    for (int q = b; q < e; ++q)
                    {
                        const int s = p.row_slot[q];
                        const uint32_t nk = ~K[s];
                        two |= one & nk;
                        one |= nk;
                        if (GEN && wp) tw ^= W[s];
                    }
                    for (int q = b; q < e; ++q)
                    {
                        const int s = p.row_slot[q];
                        const uint32_t kn = ~two & (~one | ~K[s]);
                        K[s] = kn;
                        if (GEN && wp) W[s] = kn & (tw ^ W[s]);
                    }
                }
                group_barrier(gs + 1);
                // ---- variable nodes (decoder.cpp:126-167) + posterior erasure flags ----
                uint32_t any_e = 0;
                err_bits = 0;
                err_or = 0;
                for (int v = gt; v < p.nc; v += BEC_TPG)
                {
                    const int b = p.col_ptr[v], e = p.col_ptr[v + 1], vw = e - b;
                    const uint32_t rec = ~er[v]; // frames that received this bit: they send it on every edge
                    uint32_t known;
                    if (vw == 0) known = rec;
                    else if (vw == 1)
                    {
                        const int s0 = p.col_slot[b];
                        known = rec | K[s0];
                        K[s0] = p.deg1_compat ? 0xFFFFFFFFu : rec;
                        if (GEN && wp) W[s0] = ~rec & xi[v];
                    }
                    else if (vw == 2)
                    {
                        const int s0 = p.col_slot[b], s1 = p.col_slot[b + 1];
                        const uint32_t k0 = K[s0], k1 = K[s1];
                        uint32_t w0 = 0, w1 = 0;
                        if (GEN && wp) { w0 = W[s0]; w1 = W[s1]; }
                        known = rec | (k0 & ~w0) | (k1 & ~w1);
                        K[s0] = rec | k1;
                        K[s1] = rec | k0;
                        if (GEN && wp) { W[s0] = ~rec & w1; W[s1] = ~rec & w0; }
                    }
                    else
                    {
                        uint32_t one = 0, two = 0; // frames with >= 1 / >= 2 known-and-right inputs
                        for (int q = b; q < e; ++q)
                        {
                            const int s = p.col_slot[q];
                            uint32_t k = K[s];
                            if (GEN && wp) k &= ~W[s];
                            two |= one & k;
                            one |= k;
                        }
                        for (int q = b; q < e; ++q)
                        {
                            const int s = p.col_slot[q];
                            uint32_t k = K[s];
                            if (GEN && wp) { k &= ~W[s]; W[s] = 0u; }
                            K[s] = rec | two | (one & ~k);
                        }
                        known = rec | one;
                    }
                    const uint32_t erased = ~known & valid;
                    any_e |= erased;
                    if (p.tx_flag[v])
                    { // decision = true bit unless the posterior is 'E', then 1: an error iff erased and the true bit is 0
                        const uint32_t bad = has_g ? (erased & ~xi[v]) : erased;
                        err_bits += (uint32_t)__popc(bad);
                        err_or |= bad;
                        if (GEN && logging) pe[v] = bad;
                    }
                    else if (GEN && logging) pe[v] = 0u;
                }
                any_e = __reduce_or_sync(0xffffffffu, any_e);
                if (lane == 0 && any_e) atomicOr(&red[it & 1], any_e);
                group_barrier(gs + 1);
                still = red[it & 1];
                if (gt == 0) red[(it + 1) & 1] = 0; // consumed two barriers ago
                ret_sum += (uint32_t)__popc(still); // frames that have to go on: one more completed iteration in their count
                if (GEN && logging && gt < 32 && ((still >> gt) & 1u)) itc[gt] += 1;
                if (p.early_term && still == 0) { ++it; break; }
            }
            // ---- accounting (ldpcsim.cpp:178-190) ----
            err_bits = __reduce_add_sync(0xffffffffu, err_bits);
            err_or = __reduce_or_sync(0xffffffffu, err_or);
            if (lane == 0)
            {
                if (err_bits) atomicAdd(&red[2], err_bits);
                if (err_or) atomicOr(&red[3], err_or);
            }
            group_barrier(gs + 1);
            if (gt == 0)
            {
                const uint32_t nvalid = (uint32_t)__popc(valid);
                const unsigned long long ret = p.early_term ? (unsigned long long)ret_sum : (unsigned long long)nvalid * (unsigned long long)p.max_iter;
                const unsigned long long done_frames = p.early_term ? (unsigned long long)__popc(valid & ~still) : 0ull;
                atomicAdd(&s_cnt[0], (unsigned long long)__popc(red[3]));
                atomicAdd(&s_cnt[1], (unsigned long long)red[2]);
                atomicAdd(&s_cnt[2], (unsigned long long)nvalid);
                atomicAdd(&s_cnt[3], ret);
                atomicAdd(&s_cnt[4], ret + done_frames); // iterations the reference executes: the successful one included
            }
            if constexpr (GEN)
            {
                if (logging)
                { // one record per frame in error: {global frame, bit errors, reference iteration count}
                    uint32_t em = red[3];
                    while (em)
                    {
                        const int f = __ffs(em) - 1;
                        em &= em - 1;
                        uint32_t n = 0;
                        for (int v = gt; v < p.nc; v += BEC_TPG) n += (pe[v] >> f) & 1u;
                        n = __reduce_add_sync(0xffffffffu, n);
                        if (lane == 0 && n) atomicAdd(&red[4], n);
                        group_barrier(gs + 1);
                        if (gt == 0)
                        {
                            const unsigned long long slot = atomicAdd(p.err_count, 1ull);
                            if (slot < p.err_cap)
                            {
                                const uint32_t ret = p.early_term ? itc[f] : (uint32_t)p.max_iter;
                                p.err_log[2 * slot] = p.frame0 + f0 + (uint64_t)f;
                                p.err_log[2 * slot + 1] = (unsigned long long)red[4] | ((unsigned long long)ret << 32);
                            }
                            red[4] = 0;
                        }
                        group_barrier(gs + 1);
                    }
                }
            }
            group_barrier(gs + 1); // red[] is reused by the next word
        }
        __syncthreads();
        if (tid < 5 && s_cnt[tid]) atomicAdd(&p.counters[tid], s_cnt[tid]);
    }
} // namespace b200
