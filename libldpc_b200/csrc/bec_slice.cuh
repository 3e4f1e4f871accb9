// Bit-sliced erasure-channel simulation kernel: BEC channel -> erasure message passing -> error accounting for the
// sweep path (reference: channel_bec, src/sim/channel.cpp:193-229; ldpc_decoder_bec::decode, src/decoding/decoder.cpp:91-192;
// vn_update / cn_update, src/decoding/decoder.h:145-155; accounting, src/sim/ldpcsim.cpp:184-190).
//
// The sweep transmits the all-zero codeword, so every message the reference exchanges is either 'E' or the known bit 0
// (a check output is the XOR of known zeros, a variable output is the true bit): the decoder state per edge is ONE bit,
// "known".  32 frames are packed into one 32-bit word per edge (frame f = bit f), and the reference's forward/backward
// recursions collapse to word-wide boolean folds that give the same result bit for bit:
//   check  -> edge j is known  iff every OTHER input is known        (cn_update folds to "E if any E")
//   erased variable of degree >= 3 -> edge j is known iff some OTHER input is known (vn_update folds to "true bit if any
//            input equals it"); degree 2 swaps the two raw inputs, degree 1 sends 0 (the reference's out-of-bounds read,
//            SURVEY T13; 'E' with bec_deg1_compat = 0), posterior known iff any input is known
//   received variable -> known everywhere.
// Early termination (no 'E' left among the posteriors, decoder.cpp:169-186) is per frame: knowledge only grows, so a
// finished frame is a fixed point and simply rides along until its 32-frame word retires; the reference's return value
// (iterations completed BEFORE the successful one) is the number of iterations after which the frame still had erasures.
//
// One group of TPG threads owns one 32-frame word: messages live in shared memory in place (a check / a variable reads
// and rewrites only its own edges), groups synchronise on their own named barrier and pull words independently.
#pragma once
#include "kernels.cuh"

namespace b200
{
    struct BecSliceParams
    {
        const int32_t *row_ptr, *row_edge; // check -> edge ids (file order)
        const int32_t *col_ptr, *col_edge; // variable -> edge ids (file order)
        const int32_t *tx_var, *punct, *shorten;
        const uint8_t *tx_flag; // [nc] 1 = transmitted position (counts towards bit / frame errors)
        int nc, mc, nnz, nct, n_punct, n_short;
        int max_iter, early_term, deg1_compat;
        uint32_t thr;
        uint64_t seed;
        uint32_t point;
        uint64_t frame0, n_frames;
        unsigned long long *counters; // [5] fec, bec, frames, sum(ret iters), sum(executed iterations)
        int groups_per_cta;
    };

    constexpr int BEC_TPG = 128; // threads per 32-frame word

    __device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(BEC_TPG) : "memory"); }

    __global__ void __launch_bounds__(1024, 1) bec_slice_kernel(const BecSliceParams p)
    {
        extern __shared__ __align__(16) uint32_t bs_smem[];
        __shared__ unsigned long long s_cnt[5];
        const int tid = threadIdx.x, gs = tid / BEC_TPG, gt = tid % BEC_TPG, lane = tid & 31;
        const int words = p.nnz + p.nc + 8; // per group: msg[nnz] | er[nc] | red[8]
        uint32_t *msg = bs_smem + (size_t)gs * words, *er = msg + p.nnz, *red = er + p.nc;
        if (tid < 5) s_cnt[tid] = 0;
        __syncthreads();

        const uint64_t n_words = (p.n_frames + 31) / 32;
        for (uint64_t w = (uint64_t)blockIdx.x * p.groups_per_cta + gs; w < n_words; w += (uint64_t)gridDim.x * p.groups_per_cta)
        {
            const uint64_t f0 = w * 32;
            const uint32_t valid = (p.n_frames - f0 >= 32) ? 0xFFFFFFFFu : ((1u << (uint32_t)(p.n_frames - f0)) - 1u);

            // ---- channel: er[v] bit f = frame f0+f sees an erasure at variable v (channel.cpp:193-229) ----
            for (int i = gt; i < p.n_punct; i += BEC_TPG) er[p.punct[i]] = 0xFFFFFFFFu; // punctured: 'E'
            for (int i = gt; i < p.n_short; i += BEC_TPG) er[p.shorten[i]] = 0u;         // shortened: the (zero) true bit
            if (gt < 8) red[gt] = 0;
            const int nblk = (p.nct + 3) >> 2;
            for (int q = gt; q < nblk; q += BEC_TPG)
            {
                uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
#pragma unroll 4
                for (int f = 0; f < 32; ++f)
                { // same stream as the byte-wise path: counter (block q, frame), values 4q .. 4q+3
                    const u32x4 r = channel_block(p.seed, p.point, 0, p.frame0 + f0 + f, (uint32_t)q);
                    e0 |= (r.x < p.thr ? 1u : 0u) << f;
                    e1 |= (r.y < p.thr ? 1u : 0u) << f;
                    e2 |= (r.z < p.thr ? 1u : 0u) << f;
                    e3 |= (r.w < p.thr ? 1u : 0u) << f;
                }
                const int t = 4 * q;
                er[p.tx_var[t]] = e0;
                if (t + 1 < p.nct) er[p.tx_var[t + 1]] = e1;
                if (t + 2 < p.nct) er[p.tx_var[t + 2]] = e2;
                if (t + 3 < p.nct) er[p.tx_var[t + 3]] = e3;
            }
            group_barrier(gs + 1);
            // ---- v2c of iteration 0 = the channel value (decoder.cpp:96-99): known iff received ----
            for (int v = gt; v < p.nc; v += BEC_TPG)
            {
                const uint32_t k = ~er[v];
                for (int q = p.col_ptr[v]; q < p.col_ptr[v + 1]; ++q) msg[p.col_edge[q]] = k;
            }
            group_barrier(gs + 1);

            uint32_t ret_sum = 0, err_bits = 0, err_or = 0, still = valid;
            int it = 0;
            for (; it < p.max_iter; ++it)
            {
                // ---- check nodes (decoder.cpp:105-123): output known iff all other inputs are known ----
                for (int c = gt; c < p.mc; c += BEC_TPG)
                {
                    const int b = p.row_ptr[c], e = p.row_ptr[c + 1];
                    if (e - b < 2) continue;
                    uint32_t one = 0, two = 0; // frames with >= 1 / >= 2 unknown inputs
                    for (int q = b; q < e; ++q)
                    {
                        const uint32_t nk = ~msg[p.row_edge[q]];
                        two |= one & nk;
                        one |= nk;
                    }
                    for (int q = b; q < e; ++q)
                    {
                        const int ed = p.row_edge[q];
                        msg[ed] = ~two & (~one | ~msg[ed]);
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
                        const int e0 = p.col_edge[b];
                        known = rec | msg[e0];
                        msg[e0] = p.deg1_compat ? 0xFFFFFFFFu : rec;
                    }
                    else if (vw == 2)
                    {
                        const int e0 = p.col_edge[b], e1 = p.col_edge[b + 1];
                        const uint32_t c0 = msg[e0], c1 = msg[e1];
                        known = rec | c0 | c1;
                        msg[e0] = rec | c1;
                        msg[e1] = rec | c0;
                    }
                    else
                    {
                        uint32_t one = 0, two = 0; // frames with >= 1 / >= 2 known inputs
                        for (int q = b; q < e; ++q)
                        {
                            const uint32_t k = msg[p.col_edge[q]];
                            two |= one & k;
                            one |= k;
                        }
                        for (int q = b; q < e; ++q)
                        {
                            const int ed = p.col_edge[q];
                            msg[ed] = rec | two | (one & ~msg[ed]);
                        }
                        known = rec | one;
                    }
                    const uint32_t erased = ~known & valid;
                    any_e |= erased;
                    if (p.tx_flag[v])
                    { // decision = true bit unless the posterior is 'E' ("wrong bit"): errors = erased transmitted positions
                        err_bits += (uint32_t)__popc(erased);
                        err_or |= erased;
                    }
                }
                any_e = __reduce_or_sync(0xffffffffu, any_e);
                if (lane == 0 && any_e) atomicOr(&red[it & 1], any_e);
                group_barrier(gs + 1);
                still = red[it & 1];
                if (gt == 0) red[(it + 1) & 1] = 0; // consumed two barriers ago
                ret_sum += (uint32_t)__popc(still); // frames that have to go on: one more completed iteration in their count
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
            group_barrier(gs + 1); // red[] is reused by the next word
        }
        __syncthreads();
        if (tid < 5 && s_cnt[tid]) atomicAdd(&p.counters[tid], s_cnt[tid]);
    }
} // namespace b200
