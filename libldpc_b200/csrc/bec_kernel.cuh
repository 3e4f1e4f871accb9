// Erasure-channel decoder kernel (reference: ldpc_decoder_bec, src/decoding/decoder.cpp:91-192,
// vn_update / cn_update src/decoding/decoder.h:145-155).  Messages are bytes in {0, 1, 'E'}.
// Mapping: one warp lane per frame (32 frames per CTA), one warp per node; message arrays
// [slot][32] bytes live in L2/HBM.  The forward/backward recursions of the reference are executed
// literally (forward values parked in the output slots) so that every quirk of the byte-level rules
// — the genie check against the true bit, "wrong bit = 1" — is reproduced.
#pragma once
#include "kernels.cuh"

namespace b200
{
    constexpr uint8_t BEC_E = 69; // 'E', src/core/functions.h:105

    struct BecParams
    {
        const uint32_t *cn_desc, *vn_desc;
        const void *cn_col, *vn_slot, *vn_id;
        const int32_t *bit_pos, *punct, *shorten;
        int cn_rounds, vn_rounds, n_slots;
        int nc, nct, n_punct, n_short;
        int max_iter, early_term, deg1_compat;
        int kind; // SRC_LLR (buffers) or SRC_BEC (Philox)
        const uint8_t *in, *cw; // decode mode: [n][nc]
        uint32_t thr;
        uint64_t seed;
        uint32_t point;
        uint64_t frame0, n_frames;
        uint8_t *out, *hard;
        int32_t *iters_out;
        unsigned long long *counters;
        unsigned char *state;
        size_t state_stride;
        // generator matrix by column (-G: random codewords u*G, src/sim/channel.cpp:177-191); g_rows = 0 -> all-zero word
        const int32_t *g_col_ptr, *g_row;
        int g_rows, g_cols;
        // per-error diagnostics log (may be null): {global frame, bit errors | iterations << 32}
        unsigned long long *err_log, *err_count;
        unsigned long long err_cap;
    };

    __device__ __forceinline__ uint8_t bec_cn(uint8_t l, uint8_t r) { return (l == BEC_E || r == BEC_E) ? BEC_E : (uint8_t)((l != 0) ^ (r != 0)); }
    __device__ __forceinline__ uint8_t bec_vn(uint8_t l, uint8_t r, uint8_t xi) { return (xi == l || xi == r) ? xi : BEC_E; }

    template <typename IdxT>
    __global__ void __launch_bounds__(1024, 1) bec_kernel(const BecParams p)
    {
        __shared__ unsigned long long s_frame[32];
        __shared__ unsigned long long s_cnt[5];
        __shared__ uint32_t s_eras[32], s_err[32], s_newstate[32];
        __shared__ int s_ret[32];
        __shared__ uint32_t s_done_mask, s_next;

        const int tid = threadIdx.x, nthreads = blockDim.x, f = tid & 31, nth = tid >> 5, NT = nthreads >> 5;
        const int nc = p.nc;
        unsigned char *q = p.state + p.state_stride * blockIdx.x;
        uint8_t *v2c = q; q += (size_t)p.n_slots * 32;
        uint8_t *c2v = q; q += (size_t)p.n_slots * 32;
        uint8_t *in = q; q += (size_t)nc * 32;
        uint8_t *cw = q; q += (size_t)nc * 32;
        uint8_t *out = q;
        const IdxT *cn_col = static_cast<const IdxT *>(p.cn_col);
        const IdxT *vn_slot = static_cast<const IdxT *>(p.vn_slot);
        const IdxT *vn_id = static_cast<const IdxT *>(p.vn_id);

        if (tid < 5) s_cnt[tid] = 0;
        if (tid < 32) { s_eras[tid] = 0; s_err[tid] = 0; s_newstate[tid] = 0; }
        if (tid == 0) { s_next = 0; s_done_mask = 0; }
        __syncthreads();
        int it = 0;
        uint32_t st = 0;

        auto generate = [&](int g, unsigned long long gf)
        {
            if (p.kind == SRC_LLR)
            {
                const uint8_t *si = p.in + (size_t)gf * nc, *sc = p.cw + (size_t)gf * nc;
                for (int i = tid; i < nc; i += nthreads) { in[i * 32 + g] = si[i]; cw[i * 32 + g] = sc[i] & 1; }
                return;
            }
            const unsigned long long frame = p.frame0 + gf;
            // transmitted word: all-zero, or u*G with the information word of Philox stream 1 (same rule as the other kernels)
            for (int i = tid; i < nc; i += nthreads)
            {
                uint32_t b = 0;
                if (p.g_rows > 0 && i < p.g_cols)
                    for (int q2 = p.g_col_ptr[i]; q2 < p.g_col_ptr[i + 1]; ++q2)
                    {
                        const int r = p.g_row[q2];
                        const u32x4 w = channel_block(p.seed, p.point, 1, frame, (uint32_t)r >> 7);
                        const uint32_t q4[4] = {w.x, w.y, w.z, w.w};
                        b ^= q4[(r >> 5) & 3] >> (r & 31);
                    }
                cw[i * 32 + g] = (uint8_t)(b & 1u);
            }
            __syncthreads();
            const int nblk = (p.nct + 3) >> 2;
            for (int j = tid; j < nblk; j += nthreads)
            { // y = erasure w.p. eps else x (src/sim/channel.cpp:193-205)
                const u32x4 r = channel_block(p.seed, p.point, 0, frame, (uint32_t)j);
                const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const int t = j + k * nblk; // value k of block j <-> transmitted index j + k*nblk (channel specification)
                    if (t < p.nct) in[p.bit_pos[t] * 32 + g] = (w[k] < p.thr) ? BEC_E : cw[p.bit_pos[t] * 32 + g];
                }
            }
            for (int i = tid; i < p.n_punct; i += nthreads) in[p.punct[i] * 32 + g] = BEC_E; // channel.cpp:212-215
            for (int i = tid; i < p.n_short; i += nthreads) in[p.shorten[i] * 32 + g] = cw[p.shorten[i] * 32 + g]; // the true bit, channel.cpp:219-222
        };

        auto retire_and_refill = [&](uint32_t mask, bool write_outputs)
        {
            if (write_outputs && (p.out || p.hard || p.iters_out))
            {
                for (int g = 0; g < 32; ++g)
                    if ((mask >> g) & 1u)
                    {
                        const size_t o = (size_t)s_frame[g] * nc;
                        for (int i = tid; i < nc; i += nthreads)
                        {
                            const uint8_t v = out[i * 32 + g];
                            if (p.out) p.out[o + i] = v;
                            if (p.hard) p.hard[o + i] = (v == BEC_E) ? (uint8_t)1 : cw[i * 32 + g]; // decoder.cpp:165
                        }
                        if (tid == 0 && p.iters_out) p.iters_out[s_frame[g]] = s_ret[g];
                    }
                __syncthreads();
            }
            if (tid == 0)
                for (int g = 0; g < 32; ++g)
                    if ((mask >> g) & 1u)
                    {
                        const unsigned long long gf = (unsigned long long)blockIdx.x + (unsigned long long)gridDim.x * s_next;
                        if (gf < p.n_frames) { s_frame[g] = gf; s_newstate[g] = 1; ++s_next; }
                        else s_newstate[g] = 0;
                    }
            __syncthreads();
            for (int g = 0; g < 32; ++g)
                if (((mask >> g) & 1u) && s_newstate[g]) generate(g, s_frame[g]);
            if ((mask >> f) & 1u) { st = s_newstate[f]; it = 0; }
        };

        retire_and_refill(0xFFFFFFFFu, false);

        for (;;)
        {
            if (!__syncthreads_or(st != 0)) break;
            // ---- check nodes: F/B recursion of decoder.cpp:105-123 ---------------------------------
            if (st)
            {
                const bool first = (it == 0);
                for (int r = 0; r < p.cn_rounds; ++r)
                {
                    const uint32_t d = p.cn_desc[r * NT + nth];
                    if (d == IDLE) continue;
                    const uint32_t p0 = d & 0xFFFFFFu;
                    const int deg = (int)(d >> 24);
                    auto v = [&](int k) -> uint8_t { return first ? in[(uint32_t)cn_col[p0 + k] * 32 + f] : v2c[(p0 + k) * 32 + f]; };
                    uint8_t Fp = v(0);
                    for (int k = 1; k < deg; ++k) { c2v[(p0 + k) * 32 + f] = Fp; Fp = bec_cn(Fp, v(k)); }
                    uint8_t B = v(deg - 1);
                    for (int k = deg - 2; k >= 1; --k)
                    {
                        const uint32_t idx = (p0 + k) * 32 + f;
                        c2v[idx] = bec_cn(c2v[idx], B);
                        B = bec_cn(B, v(k));
                    }
                    c2v[p0 * 32 + f] = B;
                }
            }
            __syncthreads();
            // ---- variable nodes: decoder.cpp:126-167 ------------------------------------------------
            uint32_t err = 0, eras = 0;
            if (st)
            {
                for (int r = 0; r < p.vn_rounds; ++r)
                {
                    const uint32_t d = p.vn_desc[r * NT + nth];
                    if (d == IDLE) continue;
                    const uint32_t id = vn_id[r * NT + nth];
                    const uint32_t q0 = d & 0x7FFFFFu;
                    const int deg = (int)((d >> 23) & 0xFFu);
                    const uint8_t xi = cw[id * 32 + f];
                    uint8_t o;
                    if (in[id * 32 + f] != BEC_E)
                    {
                        for (int k = 0; k < deg; ++k) v2c[(uint32_t)vn_slot[q0 + k] * 32 + f] = xi;
                        o = xi;
                    }
                    else if (deg == 0) o = BEC_E;
                    else if (deg == 1)
                    { // the reference reads one element before its scratch vector here (decoder.cpp:155-156)
                        const uint32_t s0 = (uint32_t)vn_slot[q0] * 32 + f;
                        v2c[s0] = p.deg1_compat ? (uint8_t)0 : BEC_E;
                        o = c2v[s0];
                    }
                    else
                    {
                        auto c = [&](int k) -> uint8_t { return c2v[(uint32_t)vn_slot[q0 + k] * 32 + f]; };
                        uint8_t Fp = c(0);
                        for (int k = 1; k < deg; ++k) { v2c[(uint32_t)vn_slot[q0 + k] * 32 + f] = Fp; Fp = bec_vn(Fp, c(k), xi); }
                        o = Fp; // mExMsgF[vw-1]
                        uint8_t B = c(deg - 1);
                        for (int k = deg - 2; k >= 1; --k)
                        {
                            const uint32_t idx = (uint32_t)vn_slot[q0 + k] * 32 + f;
                            v2c[idx] = bec_vn(v2c[idx], B, xi);
                            B = bec_vn(B, c(k), xi);
                        }
                        v2c[(uint32_t)vn_slot[q0] * 32 + f] = B;
                    }
                    out[id * 32 + f] = o;
                    eras |= (o == BEC_E) ? 1u : 0u;
                    err += ((d >> 31) && o == BEC_E && xi == 0) ? 1u : 0u; // estimate 1 ("wrong bit", gf2.cpp:5-8) vs true bit
                }
                ++it;
                if (eras) s_eras[f] = 1;
                if (err) atomicAdd(&s_err[f], err);
            }
            __syncthreads();
            if (tid < 32)
            {
                bool fin = false;
                if (st)
                {
                    const bool clean = p.early_term && s_eras[f] == 0; // decoder.cpp:169-186
                    if (clean || it >= p.max_iter)
                    {
                        const int ret = clean ? it - 1 : p.max_iter;
                        const uint32_t e = s_err[f];
                        atomicAdd(&s_cnt[0], (unsigned long long)(e ? 1 : 0));
                        atomicAdd(&s_cnt[1], (unsigned long long)e);
                        atomicAdd(&s_cnt[2], 1ull);
                        atomicAdd(&s_cnt[3], (unsigned long long)ret);
                        atomicAdd(&s_cnt[4], (unsigned long long)it);
                        s_ret[f] = ret;
                        if (e && p.err_log)
                        {
                            const unsigned long long slot = atomicAdd(p.err_count, 1ull);
                            if (slot < p.err_cap)
                            {
                                p.err_log[2 * slot] = p.frame0 + s_frame[f];
                                p.err_log[2 * slot + 1] = (unsigned long long)e | ((unsigned long long)(uint32_t)ret << 32);
                            }
                        }
                        fin = true;
                    }
                }
                s_eras[f] = 0;
                s_err[f] = 0;
                const uint32_t m = __ballot_sync(0xffffffffu, fin);
                if (tid == 0) s_done_mask = m;
            }
            __syncthreads();
            const uint32_t dm = s_done_mask;
            if (dm) retire_and_refill(dm, true);
        }
        __syncthreads();
        if (tid < 5 && s_cnt[tid]) atomicAdd(&p.counters[tid], s_cnt[tid]);
    }
} // namespace b200
