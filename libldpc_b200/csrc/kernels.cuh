// sm_100a kernels of the decode hot path: Philox channel -> flooding min-sum / box-plus decode ->
// hard decision + error accounting, fused into ONE persistent kernel.
//
// Mapping ("tile"): a CTA owns `fpc` frame lanes; thread t serves frame lane f = t % fpc as node
// thread t / fpc.  All per-frame arrays are laid out [index][fpc] with the frame lane fastest, so a
// warp touching npw = 32/fpc consecutive indices moves one contiguous 32*sizeof(T) block (conflict
// free in shared memory, fully coalesced in HBM).  State per lane: c2v message per edge slot, the
// posterior `out` per variable and the channel LLR per variable; v2c is never stored — it is
// recomputed as out - c2v, which is exactly the value the reference stores
// (src/decoding/decoder.cpp:60-63), so results stay bit-identical while one of the reference's two
// message arrays disappears.
//
// A lane that finishes its frame (syndrome clear after an iteration, or iteration limit) is refilled
// at once with the next frame (LLRs regenerated from the counter-based Philox stream), so early
// termination never leaves lanes idle waiting for the slowest frame of a batch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200
{
    // ------------------------------------------------------------------------------------------
    // Philox4x32-10 (Salmon et al., SC'11) — counter-based, stateless
    // ------------------------------------------------------------------------------------------
    struct u32x4 { uint32_t x, y, z, w; };

    __host__ __device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1)
    {
#pragma unroll
        for (int r = 0; r < 10; ++r)
        {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
            const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
            u32x4 n;
            n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
            n.y = (uint32_t)p1;
            n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
            n.w = (uint32_t)p0;
            c = n;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        return c;
    }

    // counter = (block index j, frame lo, frame hi, point | stream << 24), key = seed
    __host__ __device__ __forceinline__ u32x4 channel_block(uint64_t seed, uint32_t point, uint32_t stream, uint64_t frame, uint32_t j)
    {
        u32x4 c = {j, (uint32_t)frame, (uint32_t)(frame >> 32), (point & 0xFFFFFFu) | (stream << 24)};
        return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    }

    enum { SRC_LLR = 0, SRC_AWGN = 1, SRC_BSC = 2, SRC_BEC = 3 };
    enum { ALG_MS = 0, ALG_BP = 1 };
    constexpr uint32_t IDLE = 0xFFFFFFFFu;

    struct KParams
    {
        // code tables (device global memory)
        const uint32_t *cn_desc, *vn_desc;
        const void *cn_col, *vn_slot, *vn_id; // IdxT arrays
        const int32_t *bit_pos, *punct, *shorten;
        int cn_rounds, vn_rounds, n_slots, n_vslots;
        int nc, nct, n_punct, n_short;
        int fpc, fshift;
        // decoder
        int max_iter, early_term;
        // frame source
        int kind;
        const double *llr_in; // SRC_LLR: [n_frames][nc]
        double sigma, sigma2, delta;
        uint32_t thr;
        uint64_t seed;
        uint32_t point;
        uint64_t frame0, n_frames;
        // sinks (indexed by frame - frame0); any may be null
        double *llr_out;
        uint8_t *hard_out;
        int32_t *iters_out;
        unsigned long long *counters; // [5] fec, bec, frames, sum(ret iters), sum(executed iters)
        // global-memory residency: per-CTA state block
        unsigned char *state;
        size_t state_stride;
    };

    template <typename T> struct Num;
    template <> struct Num<double>
    {
        static __device__ __forceinline__ uint32_t sign(double v) { return (uint32_t)__double2hiint(v) >> 31; }
        static __device__ __forceinline__ double abs(double v) { return fabs(v); }
        static __device__ __forceinline__ double with_sign(double mag, uint32_t s) { return __hiloint2double(__double2hiint(mag) | (int)(s << 31), __double2loint(mag)); }
        static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
        static __device__ __forceinline__ double exp_(double v) { return exp(v); }
        static __device__ __forceinline__ double log_(double v) { return log(v); }
    };
    template <> struct Num<float>
    {
        static __device__ __forceinline__ uint32_t sign(float v) { return __float_as_uint(v) >> 31; }
        static __device__ __forceinline__ float abs(float v) { return fabsf(v); }
        static __device__ __forceinline__ float with_sign(float mag, uint32_t s) { return __uint_as_float(__float_as_uint(mag) | (s << 31)); }
        static __device__ __forceinline__ float inf() { return __uint_as_float(0x7f800000u); }
        static __device__ __forceinline__ float exp_(float v) { return __expf(v); }
        static __device__ __forceinline__ float log_(float v) { return __logf(v); }
    };

    // ------------------------------------------------------------------------------------------
    // check-node updates.  slot(k) = p0 + k*npw, value index = slot*fpc + f.
    // Both return the parity of the hard decisions of the check's variables (the syndrome bit of the
    // previous iteration's output, reference: decoder.h:47-64) — computed from the gathered `out`
    // values for free.
    // ------------------------------------------------------------------------------------------

    // min-sum: the forward/backward recursion of decoder.cpp:30-44 with f = minsum (decoder.h:17-20)
    // yields, for every edge, (product of the other signs) * (minimum of the other magnitudes); with
    // sign *bits* (std::signbit, -0.0 negative) that is reproduced exactly by min1/min2 + sign parity.
    template <typename T, typename IdxT>
    __device__ __forceinline__ uint32_t cn_minsum(const T *__restrict__ src, T *__restrict__ c2v, const IdxT *__restrict__ cn_col,
                                                  uint32_t p0, int deg, int npw, int fpc, int f, bool first)
    {
        T min1 = Num<T>::inf(), min2 = Num<T>::inf();
        int arg = 0;
        unsigned long long smask = 0;
        uint32_t par = 0;
#pragma unroll 4
        for (int k = 0; k < deg; ++k)
        {
            const uint32_t slot = p0 + k * npw;
            const uint32_t col = cn_col[slot];
            const T o = src[col * fpc + f];
            const T c = first ? T(0) : c2v[slot * fpc + f];
            const T v = o - c; // == the reference's stored v2c (decoder.cpp:62), or LLRin on the first pass (:18)
            par ^= (o <= T(0)) ? 1u : 0u;
            smask |= (unsigned long long)Num<T>::sign(v) << k;
            const T a = Num<T>::abs(v);
            const bool lt1 = a < min1, lt2 = a < min2;
            min2 = lt1 ? min1 : (lt2 ? a : min2);
            arg = lt1 ? k : arg;
            min1 = lt1 ? a : min1;
        }
        const uint32_t tot = (uint32_t)__popcll(smask) & 1u;
#pragma unroll 4
        for (int k = 0; k < deg; ++k)
        {
            const uint32_t slot = p0 + k * npw;
            const T mag = (k == arg) ? min2 : min1;
            const uint32_t s = tot ^ (uint32_t)((smask >> k) & 1ull);
            c2v[slot * fpc + f] = Num<T>::with_sign(mag, s);
        }
        return par;
    }

    // pairwise box-plus with the Jacobian correction, decoder.h:12-15 (same expression, same order)
    template <typename T>
    __device__ __forceinline__ T boxplus(T x, T y)
    {
        const T ax = Num<T>::abs(x), ay = Num<T>::abs(y);
        const T m = (ay < ax) ? ay : ax;
        const T sm = Num<T>::with_sign(m, Num<T>::sign(x) ^ Num<T>::sign(y));
        const T num = T(1) + Num<T>::exp_(-Num<T>::abs(x + y));
        const T den = T(1) + Num<T>::exp_(-Num<T>::abs(x - y));
        return sm + Num<T>::log_(num / den);
    }

    // sum-product by the reference's forward/backward recursion (decoder.cpp:30-44), file order.
    template <typename T, typename IdxT, int MAXD>
    __device__ __forceinline__ uint32_t cn_boxplus(const T *__restrict__ src, T *__restrict__ c2v, const IdxT *__restrict__ cn_col,
                                                   uint32_t p0, int deg, int npw, int fpc, int f, bool first)
    {
        T v[MAXD], F[MAXD];
        uint32_t par = 0;
#pragma unroll
        for (int k = 0; k < MAXD; ++k)
            if (k < deg)
            {
                const uint32_t slot = p0 + k * npw;
                const T o = src[(uint32_t)cn_col[slot] * fpc + f];
                const T c = first ? T(0) : c2v[slot * fpc + f];
                v[k] = o - c;
                par ^= (o <= T(0)) ? 1u : 0u;
            }
        F[0] = v[0];
#pragma unroll
        for (int k = 1; k < MAXD; ++k)
            if (k < deg) F[k] = boxplus(F[k - 1], v[k]);
        // backward sweep: B holds the combination of v[k+1..deg-1]
        T B = T(0);
#pragma unroll
        for (int k = MAXD - 1; k >= 0; --k)
            if (k < deg)
            {
                T r;
                if (k == deg - 1) { r = F[(k > 0) ? k - 1 : 0]; B = v[k]; }                     // c2v[last] = F[cw-2]
                else if (k == 0) { r = B; }                                                     // c2v[0] = B[1]
                else { r = boxplus(F[k - 1], B); B = boxplus(B, v[k]); }                         // f(F[j-1], B[j+1])
                c2v[(p0 + k * npw) * fpc + f] = r;
            }
        return par;
    }

    // same recursion for arbitrary degree: F is parked in the (about to be overwritten) message slots
    template <typename T, typename IdxT>
    __device__ __noinline__ uint32_t cn_boxplus_any(const T *__restrict__ src, T *__restrict__ c2v, const IdxT *__restrict__ cn_col,
                                                    uint32_t p0, int deg, int npw, int fpc, int f, bool first)
    {
        T v[64];
        uint32_t par = 0;
        for (int k = 0; k < deg; ++k)
        {
            const uint32_t slot = p0 + k * npw;
            const T o = src[(uint32_t)cn_col[slot] * fpc + f];
            const T c = first ? T(0) : c2v[slot * fpc + f];
            v[k] = o - c;
            par ^= (o <= T(0)) ? 1u : 0u;
        }
        T Fp = v[0]; // F[k-1] while visiting k
        for (int k = 1; k < deg; ++k)
        {
            c2v[(p0 + k * npw) * fpc + f] = Fp; // park F[k-1] in slot k
            Fp = boxplus(Fp, v[k]);
        }
        T B = v[deg - 1]; // slot deg-1 already holds F[deg-2] == its final value
        for (int k = deg - 2; k >= 1; --k)
        {
            const uint32_t idx = (p0 + k * npw) * fpc + f;
            c2v[idx] = boxplus(c2v[idx], B);
            B = boxplus(B, v[k]);
        }
        c2v[p0 * fpc + f] = B;
        return par;
    }

    // ------------------------------------------------------------------------------------------
    // the persistent tile kernel
    // ------------------------------------------------------------------------------------------
    template <typename T, typename IdxT, int ALG, bool SMEM>
    __global__ void __launch_bounds__(ALG == ALG_MS ? 1024 : 512, 1) tile_kernel(const KParams p)
    {
        extern __shared__ __align__(16) unsigned char dyn_smem[];
        __shared__ unsigned long long s_frame[32];
        __shared__ unsigned long long s_cnt[5];
        __shared__ uint32_t s_synd[32], s_err[32], s_newstate[32];
        __shared__ int s_ret[32];
        __shared__ uint32_t s_done_mask, s_next;

        const int tid = threadIdx.x, nthreads = blockDim.x;
        const int fpc = p.fpc;
        const int f = tid & (fpc - 1);
        const int nth = tid >> p.fshift;     // node thread
        const int NT = nthreads >> p.fshift; // node threads per CTA
        const int npw = 32 >> p.fshift;
        const int lane = tid & 31;
        const int nc = p.nc;

        // ---- carve state and tables --------------------------------------------------------
        T *c2v, *out, *llr;
        const uint32_t *cn_desc, *vn_desc;
        const IdxT *cn_col, *vn_slot, *vn_id;
        if (SMEM)
        {
            unsigned char *q = dyn_smem;
            c2v = reinterpret_cast<T *>(q); q += sizeof(T) * (size_t)p.n_slots * fpc;
            out = reinterpret_cast<T *>(q); q += sizeof(T) * (size_t)nc * fpc;
            llr = reinterpret_cast<T *>(q); q += sizeof(T) * (size_t)nc * fpc;
            uint32_t *cd = reinterpret_cast<uint32_t *>(q); q += 4 * (size_t)p.cn_rounds * NT;
            uint32_t *vd = reinterpret_cast<uint32_t *>(q); q += 4 * (size_t)p.vn_rounds * NT;
            IdxT *cc = reinterpret_cast<IdxT *>(q); q += sizeof(IdxT) * (size_t)p.n_slots;
            IdxT *vs = reinterpret_cast<IdxT *>(q); q += sizeof(IdxT) * (size_t)p.n_vslots;
            IdxT *vi = reinterpret_cast<IdxT *>(q);
            for (int i = tid; i < p.cn_rounds * NT; i += nthreads) cd[i] = p.cn_desc[i];
            for (int i = tid; i < p.vn_rounds * NT; i += nthreads) { vd[i] = p.vn_desc[i]; vi[i] = static_cast<const IdxT *>(p.vn_id)[i]; }
            for (int i = tid; i < p.n_slots; i += nthreads) cc[i] = static_cast<const IdxT *>(p.cn_col)[i];
            for (int i = tid; i < p.n_vslots; i += nthreads) vs[i] = static_cast<const IdxT *>(p.vn_slot)[i];
            cn_desc = cd; vn_desc = vd; cn_col = cc; vn_slot = vs; vn_id = vi;
        }
        else
        {
            unsigned char *q = p.state + p.state_stride * blockIdx.x;
            c2v = reinterpret_cast<T *>(q); q += sizeof(T) * (size_t)p.n_slots * fpc;
            out = reinterpret_cast<T *>(q); q += sizeof(T) * (size_t)nc * fpc;
            llr = reinterpret_cast<T *>(q);
            cn_desc = p.cn_desc; vn_desc = p.vn_desc;
            cn_col = static_cast<const IdxT *>(p.cn_col);
            vn_slot = static_cast<const IdxT *>(p.vn_slot);
            vn_id = static_cast<const IdxT *>(p.vn_id);
        }
        if (tid < 5) s_cnt[tid] = 0;
        if (tid < 32) { s_synd[tid] = 0; s_err[tid] = 0; s_newstate[tid] = 0; }
        if (tid == 0) { s_next = 0; s_done_mask = 0; }
        __syncthreads();

        // replicated per-lane state machine (identical in every thread of a frame lane)
        int it = 0;      // completed iterations (variable-node phases) of the current frame
        uint32_t st = 0; // 0 idle, 1 active, 2 active but refilled after this step's check phase (skips the VN phase)

        // Writes the decoder input of global frame gf into lane g (all threads of the CTA cooperate).
        auto generate = [&](int g, unsigned long long gf)
        {
            if (p.kind == SRC_LLR)
            {
                const double *src = p.llr_in + (size_t)gf * nc;
                for (int i = tid; i < nc; i += nthreads) llr[i * fpc + g] = (T)src[i];
                return;
            }
            const unsigned long long frame = p.frame0 + gf;
            if (p.kind == SRC_AWGN)
            { // y = sigma*z + 1 (all-zero codeword, BPSK +1), LLR = 2y/sigma^2 (src/sim/channel.cpp:62-68,88-92)
                const int npairs = (p.nct + 1) >> 1;
                for (int j = tid; j < npairs; j += nthreads)
                {
                    const u32x4 r = channel_block(p.seed, p.point, 0, frame, (uint32_t)j);
                    const double u1 = ((double)((((uint64_t)r.y << 32) | r.x) >> 11) + 1.0) * 0x1p-53;
                    const double u2 = (double)((((uint64_t)r.w << 32) | r.z) >> 11) * 0x1p-53;
                    const double rad = sqrt(-2.0 * log(u1));
                    double sn, cs;
                    sincos(6.283185307179586 * u2, &sn, &cs);
                    const double y0 = __dadd_rn(__dmul_rn(rad * cs, p.sigma), 1.0);
                    const double y1 = __dadd_rn(__dmul_rn(rad * sn, p.sigma), 1.0);
                    const int t = 2 * j;
                    llr[p.bit_pos[t] * fpc + g] = (T)(__dmul_rn(2.0, y0) / p.sigma2);
                    if (t + 1 < p.nct) llr[p.bit_pos[t + 1] * fpc + g] = (T)(__dmul_rn(2.0, y1) / p.sigma2);
                }
                for (int i = tid; i < p.n_punct; i += nthreads) llr[p.punct[i] * fpc + g] = T(0);
                for (int i = tid; i < p.n_short; i += nthreads) llr[p.shorten[i] * fpc + g] = (T)99999.9;
            }
            else
            { // BSC: y = x ^ Bernoulli(eps), LLR = delta*(1-2y) (src/sim/channel.cpp:123-162)
                const int nblk = (p.nct + 3) >> 2;
                for (int j = tid; j < nblk; j += nthreads)
                {
                    const u32x4 r = channel_block(p.seed, p.point, 0, frame, (uint32_t)j);
                    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                    {
                        const int t = 4 * j + q;
                        if (t < p.nct) llr[p.bit_pos[t] * fpc + g] = (T)((w[q] < p.thr) ? -p.delta : p.delta);
                    }
                }
                for (int i = tid; i < p.n_punct; i += nthreads) llr[p.punct[i] * fpc + g] = T(0);
                for (int i = tid; i < p.n_short; i += nthreads) llr[p.shorten[i] * fpc + g] = (T)p.delta;
            }
        };

        // Retires the lanes in `mask` (results + accounting were recorded by the caller in s_ret /
        // s_cnt), hands each a new frame if any is left and regenerates its input.
        auto retire_and_refill = [&](uint32_t mask, uint32_t new_state, bool write_outputs)
        {
            if (write_outputs && (p.llr_out || p.hard_out || p.iters_out))
            {
                for (int g = 0; g < fpc; ++g)
                    if ((mask >> g) & 1u)
                    {
                        const size_t o = (size_t)s_frame[g] * nc;
                        for (int i = tid; i < nc; i += nthreads)
                        {
                            const T v = out[i * fpc + g];
                            if (p.llr_out) p.llr_out[o + i] = (double)v;
                            if (p.hard_out) p.hard_out[o + i] = (v <= T(0)) ? 1 : 0; // decoder.cpp:58
                        }
                        if (tid == 0 && p.iters_out) p.iters_out[s_frame[g]] = s_ret[g];
                    }
                __syncthreads();
            }
            if (tid == 0)
            {
                for (int g = 0; g < fpc; ++g)
                    if ((mask >> g) & 1u)
                    {
                        const unsigned long long gf = (unsigned long long)blockIdx.x + (unsigned long long)gridDim.x * s_next;
                        if (gf < p.n_frames) { s_frame[g] = gf; s_newstate[g] = new_state; ++s_next; }
                        else s_newstate[g] = 0;
                    }
            }
            __syncthreads();
            for (int g = 0; g < fpc; ++g)
                if (((mask >> g) & 1u) && s_newstate[g]) generate(g, s_frame[g]);
            if ((mask >> f) & 1u) { st = s_newstate[f]; it = 0; }
        };

        retire_and_refill(fpc == 32 ? 0xFFFFFFFFu : ((1u << fpc) - 1u), 1u, false);

        // bit pattern of the warp lanes that serve frame lane 0
        uint32_t lane_pattern = 0;
        for (int b = 0; b < 32; b += fpc) lane_pattern |= 1u << b;

        for (;;)
        {
            if (!__syncthreads_or(st != 0)) break; // also orders VN/fill writes before the check phase

            // ---- frames that reached the iteration limit without early termination retire here,
            //      before any further work is spent on them
            if (!p.early_term)
            {
                if (tid < 32)
                {
                    bool fin = false;
                    if (tid < fpc && st == 1 && it >= p.max_iter)
                    {
                        const uint32_t e = s_err[f];
                        atomicAdd(&s_cnt[0], (unsigned long long)(e ? 1 : 0));
                        atomicAdd(&s_cnt[1], (unsigned long long)e);
                        atomicAdd(&s_cnt[2], 1ull);
                        atomicAdd(&s_cnt[3], (unsigned long long)p.max_iter);
                        atomicAdd(&s_cnt[4], (unsigned long long)it);
                        s_ret[f] = p.max_iter;
                        fin = true;
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, fin);
                    if (tid == 0) s_done_mask = m;
                }
                __syncthreads();
                const uint32_t dm = s_done_mask;
                if (dm)
                {
                    retire_and_refill(dm, 1u, true);
                    if (!__syncthreads_or(st != 0)) break;
                }
            }

            // ---- check-node phase (+ syndrome of the previous iteration's decisions) ----------
            uint32_t par = 0;
            if (st)
            {
                const bool first = (it == 0);
                const T *src = first ? llr : out;
                for (int r = 0; r < p.cn_rounds; ++r)
                {
                    const uint32_t d = cn_desc[r * NT + nth];
                    if (d == IDLE) continue;
                    const uint32_t p0 = d & 0xFFFFFFu;
                    const int deg = (int)(d >> 24);
                    if (ALG == ALG_MS) par |= cn_minsum<T, IdxT>(src, c2v, cn_col, p0, deg, npw, fpc, f, first);
                    else if (deg <= 8) par |= cn_boxplus<T, IdxT, 8>(src, c2v, cn_col, p0, deg, npw, fpc, f, first);
                    else par |= cn_boxplus_any<T, IdxT>(src, c2v, cn_col, p0, deg, npw, fpc, f, first);
                }
            }
            // warp-ballot syndrome test: one vote per thread, folded per frame lane
            {
                const uint32_t b = __ballot_sync(0xffffffffu, par != 0);
                if (lane < fpc && (b & (lane_pattern << lane))) s_synd[lane] = 1;
            }
            __syncthreads();

            // ---- decision: converged (reference: decoder.cpp:66-72) or out of iterations --------
            if (tid < 32)
            {
                bool fin = false;
                if (tid < fpc)
                {
                    if (st == 1)
                    {
                        const bool conv = p.early_term && it >= 1 && s_synd[f] == 0;
                        if (conv || it >= p.max_iter)
                        {
                            const int ret = conv ? it - 1 : p.max_iter; // break happens before ++I
                            const uint32_t e = s_err[f];
                            atomicAdd(&s_cnt[0], (unsigned long long)(e ? 1 : 0));
                            atomicAdd(&s_cnt[1], (unsigned long long)e);
                            atomicAdd(&s_cnt[2], 1ull);
                            atomicAdd(&s_cnt[3], (unsigned long long)ret);
                            atomicAdd(&s_cnt[4], (unsigned long long)it);
                            s_ret[f] = ret;
                            fin = true;
                        }
                    }
                    s_synd[f] = 0;
                    s_err[f] = 0;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, fin);
                if (tid == 0) s_done_mask = m;
            }
            __syncthreads();
            {
                const uint32_t dm = s_done_mask;
                if (dm) retire_and_refill(dm, 2u, true);
            }

            // ---- variable-node phase: posterior, hard decision, bit errors ----------------------
            uint32_t err = 0;
            if (st == 1)
            {
                for (int r = 0; r < p.vn_rounds; ++r)
                {
                    const uint32_t d = vn_desc[r * NT + nth];
                    if (d == IDLE) continue;
                    const uint32_t id = vn_id[r * NT + nth];
                    const uint32_t q0 = d & 0x7FFFFFu;
                    const int deg = (int)((d >> 23) & 0xFFu);
                    T acc = llr[id * fpc + f]; // decoder.cpp:50
#pragma unroll 4
                    for (int k = 0; k < deg; ++k) acc += c2v[(uint32_t)vn_slot[q0 + k * npw] * fpc + f]; // file order, decoder.cpp:53-56
                    out[id * fpc + f] = acc;
                    err += ((d >> 31) && acc <= T(0)) ? 1u : 0u; // all-zero codeword: ldpcsim.cpp:184-188
                }
                ++it;
            }
            else if (st == 2) st = 1;
            for (int o = fpc; o < 32; o <<= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
            if (lane < fpc && err) atomicAdd(&s_err[lane], err);
        }

        __syncthreads();
        if (tid < 5 && s_cnt[tid]) atomicAdd(&p.counters[tid], s_cnt[tid]);
    }
} // namespace b200
