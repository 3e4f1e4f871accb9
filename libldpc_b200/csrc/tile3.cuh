// sm_100a decode kernel, vector-tile mapping: Philox channel -> flooding min-sum / box-plus decode ->
// hard decision + error accounting, fused into ONE persistent kernel.
//
// Mapping.  A CTA keeps FPC = LANES * VEC frames in flight (VEC = 16/sizeof(T): 2 doubles or 4
// floats).  Every per-frame array is stored as 16-byte vectors of VEC adjacent frame lanes,
// [index][LANES] vectors per index.  One thread owns ONE vector of one node: warp lane l serves node
// l / LANES of the warp's current task and vector l % LANES, so a warp walking the NPW = 32/LANES
// nodes of a task moves 512 contiguous bytes per access (conflict free in shared memory, whole
// sectors in HBM/L2) with one LDS.128/STS.128 (LDG/STG.128) per VEC messages, and every index
// load / address computation is shared by VEC frames.
//
// Work list.  Nodes of equal degree are packed NPW at a time into warp tasks, spread longest-first
// over the warps (code.cpp, TaskLayout).  A task costs one 8-byte broadcast load; its degree is
// warp-uniform, so node updates dispatch to fully unrolled fixed-degree bodies whose loads use
// immediate offsets (edge k of a task is 512 bytes after edge k-1).
//
// State per frame lane: c2v per edge slot, posterior `out` and channel LLR per variable position.
// v2c is never stored — it is recomputed as out - c2v, which is exactly the value the reference
// stores (src/decoding/decoder.cpp:60-63), so results stay bit-identical while one of the
// reference's two message arrays disappears.  A fresh frame starts with c2v = +0 and out = LLRin,
// which makes its first check pass read LLRin exactly (x - (+0) == x for every x, -0 included),
// i.e. decoder.cpp:16-19 without a special case.
//
// A frame lane that finishes (syndrome clear after an iteration, or iteration limit) is refilled at
// once with the next frame (LLRs regenerated from the counter-based Philox stream), so early
// termination never leaves lanes idle waiting for the slowest frame of a batch.
#pragma once
#include "kernels.cuh"

namespace b200
{
    // ------------------------------------------------------------------------------------------
    // 16-byte vectors of frame lanes
    // ------------------------------------------------------------------------------------------
    template <typename T> struct Vec;
    template <> struct __align__(16) Vec<double> { static constexpr int N = 2; double e[2]; };
    template <> struct __align__(16) Vec<float> { static constexpr int N = 4; float e[4]; };

    template <bool SMEM, typename T, int OFF> struct VAcc;
    template <int OFF> struct VAcc<true, double, OFF>
    {
        static __device__ __forceinline__ Vec<double> ld(uint32_t a)
        {
            Vec<double> v;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.e[0]), "=d"(v.e[1]) : "r"(a), "n"(OFF));
            return v;
        }
        static __device__ __forceinline__ void st(uint32_t a, const Vec<double> &v)
        {
            asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(v.e[0]), "d"(v.e[1]) : "memory");
        }
    };
    template <int OFF> struct VAcc<true, float, OFF>
    {
        static __device__ __forceinline__ Vec<float> ld(uint32_t a)
        {
            Vec<float> v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.e[0]), "=f"(v.e[1]), "=f"(v.e[2]), "=f"(v.e[3]) : "r"(a), "n"(OFF));
            return v;
        }
        static __device__ __forceinline__ void st(uint32_t a, const Vec<float> &v)
        {
            asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "f"(v.e[0]), "f"(v.e[1]), "f"(v.e[2]), "f"(v.e[3]) : "memory");
        }
    };
    template <typename T, int OFF> struct VAcc<false, T, OFF>
    {
        static __device__ __forceinline__ Vec<T> ld(const unsigned char *a) { return *reinterpret_cast<const Vec<T> *>(a + OFF); }
        static __device__ __forceinline__ void st(unsigned char *a, const Vec<T> &v) { *reinterpret_cast<Vec<T> *>(a + OFF) = v; }
    };

    // 8-byte task descriptors (warp-uniform address)
    template <bool SMEM> struct TaskLd;
    template <> struct TaskLd<true>
    {
        static __device__ __forceinline__ uint2 ld(uint32_t a)
        {
            uint2 v;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
            return v;
        }
    };
    template <> struct TaskLd<false>
    {
        static __device__ __forceinline__ uint2 ld(const unsigned char *a) { return __ldg(reinterpret_cast<const uint2 *>(a)); }
    };

    struct K3Params
    {
        // code tables (device global memory), TaskLayout of code.hpp
        const uint32_t *cn_task, *vn_task;    // [rounds][warps][2]
        const void *cn_col, *vn_slot;         // IdxT arrays
        const uint32_t *var_pos;              // [nc] variable id -> position
        const int32_t *tx_pos, *punct_pos, *short_pos; // positions of transmitted (ascending id) / punctured / shortened variables
        int cn_rounds, vn_rounds, n_slots, n_vslots, n_pos;
        int nc, nct, n_punct, n_short;
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
        unsigned long long *counters; // [5] fec, bec, frames, sum(ret iters), sum(executed iterations)
        // global-memory residency: per-CTA state block
        unsigned char *state;
        size_t state_stride;
    };

    // ------------------------------------------------------------------------------------------
    // check-node updates.  A node of a task is described by
    //   out_sub : &out[0][sub]                       record stride RS = 16*LANES bytes
    //   c2v0    : &c2v[p0][lane]    slot k at + k*512 (slots of a node are NPW apart, NPW*LANES*16 = 512)
    //   col0    : &cn_col[p0 + j]   entry k at + k*IS (IS = NPW*sizeof(IdxT))
    // Both algorithms return, per frame lane of the vector, the parity of the hard decisions of the
    // check's variables (= the syndrome bit of the previous iteration's output, decoder.h:47-64),
    // which comes for free with the gather.
    // ------------------------------------------------------------------------------------------
    template <typename T, typename IdxT, bool SMEM, int LANES, int ALG, int D>
    struct CnVec
    {
        typedef typename PtrOf<SMEM>::type P;
        typedef Vec<T> V;
        static constexpr int VEC = V::N, NPW = 32 / LANES, CS = 512, IS = NPW * (int)sizeof(IdxT), RS = 16 * LANES;

        template <int K> struct Step
        {
            static __device__ __forceinline__ void load(P out_sub, P c2v0, P col0, V (&v)[D], uint32_t &par)
            {
                const uint32_t col = Acc<SMEM, IdxT, K * IS>::ld(col0);
                const V o = VAcc<SMEM, T, 0>::ld(out_sub + col * RS);
                const V c = VAcc<SMEM, T, K * CS>::ld(c2v0);
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                {
                    v[K].e[e] = o.e[e] - c.e[e]; // == the reference's stored v2c (decoder.cpp:62); LLRin on a fresh frame (:18)
                    par ^= (o.e[e] <= T(0)) ? (1u << e) : 0u;
                }
                if constexpr (K + 1 < D) Step<K + 1>::load(out_sub, c2v0, col0, v, par);
            }
            static __device__ __forceinline__ void store(P c2v0, const V (&r)[D])
            {
                VAcc<SMEM, T, K * CS>::st(c2v0, r[K]);
                if constexpr (K + 1 < D) Step<K + 1>::store(c2v0, r);
            }
            // one-pass scan for min-sum: running min1/min2/argmin + sign bits, nothing else kept
            static __device__ __forceinline__ void scan(P out_sub, P c2v0, P col0, T (&min1)[VEC], T (&min2)[VEC], uint32_t (&arg)[VEC],
                                                        uint32_t (&smask)[VEC], uint32_t &par)
            {
                const uint32_t col = Acc<SMEM, IdxT, K * IS>::ld(col0);
                const V o = VAcc<SMEM, T, 0>::ld(out_sub + col * RS);
                const V c = VAcc<SMEM, T, K * CS>::ld(c2v0);
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                {
                    const T v = o.e[e] - c.e[e];
                    par ^= (o.e[e] <= T(0)) ? (1u << e) : 0u;
                    smask[e] |= (Num<T>::hi(v) >> 31) << K;
                    const T a = Num<T>::abs(v);
                    const bool lt1 = a < min1[e], lt2 = a < min2[e];
                    min2[e] = lt1 ? min1[e] : (lt2 ? a : min2[e]);
                    arg[e] = lt1 ? (uint32_t)K : arg[e];
                    min1[e] = lt1 ? a : min1[e];
                }
                if constexpr (K + 1 < D) Step<K + 1>::scan(out_sub, c2v0, col0, min1, min2, arg, smask, par);
            }
            static __device__ __forceinline__ void emit(P c2v0, const T (&min1)[VEC], const T (&min2)[VEC], const uint32_t (&arg)[VEC],
                                                        const uint32_t (&sflip)[VEC])
            {
                V r;
#pragma unroll
                for (int e = 0; e < VEC; ++e) r.e[e] = Num<T>::with_sign((arg[e] == (uint32_t)K) ? min2[e] : min1[e], sflip[e] << (31 - K));
                VAcc<SMEM, T, K * CS>::st(c2v0, r);
                if constexpr (K + 1 < D) Step<K + 1>::emit(c2v0, min1, min2, arg, sflip);
            }
        };

        static __device__ __forceinline__ T mn(T a, T b) { return (b < a) ? b : a; }

        static __device__ __forceinline__ uint32_t run(P out_sub, P c2v0, P col0)
        {
            uint32_t par = 0;
            if constexpr (ALG == ALG_MS && D > 4)
            {
                T min1[VEC], min2[VEC];
                uint32_t arg[VEC], smask[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) { min1[e] = Num<T>::inf(); min2[e] = Num<T>::inf(); arg[e] = 0; smask[e] = 0; }
                Step<0>::scan(out_sub, c2v0, col0, min1, min2, arg, smask, par);
                // sign of message k = total sign ^ own sign: bit k of sflip
#pragma unroll
                for (int e = 0; e < VEC; ++e) smask[e] ^= (__popc(smask[e]) & 1u) ? ((D >= 32) ? 0xFFFFFFFFu : ((1u << D) - 1u)) : 0u;
                Step<0>::emit(c2v0, min1, min2, arg, smask);
                return par;
            }
            else
            {
                V v[D], r[D];
                Step<0>::load(out_sub, c2v0, col0, v, par);
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                {
                    if constexpr (ALG == ALG_MS)
                    {
                        // min-sum: f = sign*sign*min (decoder.h:17-20) through the forward/backward recursion of
                        // decoder.cpp:30-44.  Magnitude: exact minimum over the other edges; sign: XOR of sign BITS
                        // (std::signbit semantics, -0.0 is negative).
                        uint32_t sx = 0;
                        T a[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) { sx ^= Num<T>::hi(v[k].e[e]); a[k] = Num<T>::abs(v[k].e[e]); }
                        T m[D];
                        if constexpr (D == 2) { m[0] = a[1]; m[1] = a[0]; }
                        else if constexpr (D == 3) { m[0] = mn(a[1], a[2]); m[1] = mn(a[0], a[2]); m[2] = mn(a[0], a[1]); }
                        else
                        {
                            const T m01 = mn(a[0], a[1]), m23 = mn(a[2], a[3]);
                            m[0] = mn(a[1], m23); m[1] = mn(a[0], m23); m[2] = mn(m01, a[3]); m[3] = mn(m01, a[2]);
                        }
#pragma unroll
                        for (int k = 0; k < D; ++k) r[k].e[e] = Num<T>::with_sign(m[k], sx ^ Num<T>::hi(v[k].e[e]));
                    }
                    else
                    {
                        // sum-product: the reference's forward/backward box-plus recursion, file order
                        T F[D];
                        F[0] = v[0].e[e];
#pragma unroll
                        for (int k = 1; k < D; ++k) F[k] = boxplus(F[k - 1], v[k].e[e]);
                        T B = v[D - 1].e[e];
                        r[D - 1].e[e] = F[D - 2];
#pragma unroll
                        for (int k = D - 2; k >= 1; --k) { r[k].e[e] = boxplus(F[k - 1], B); B = boxplus(B, v[k].e[e]); }
                        r[0].e[e] = B;
                    }
                }
                Step<0>::store(c2v0, r);
                return par;
            }
        }
    };

    // arbitrary degree (<= 64): running min1/min2 + sign mask for min-sum, parked forward values for box-plus
    template <typename T, typename IdxT, bool SMEM, int LANES, int ALG>
    __device__ __noinline__ uint32_t cn_vec_any(typename PtrOf<SMEM>::type out_sub, typename PtrOf<SMEM>::type c2v0, typename PtrOf<SMEM>::type col0, int deg)
    {
        typedef Vec<T> V;
        constexpr int VEC = V::N, NPW = 32 / LANES, CS = 512, IS = NPW * (int)sizeof(IdxT), RS = 16 * LANES;
        uint32_t par = 0;
        if (ALG == ALG_MS)
        {
            T min1[VEC], min2[VEC];
            int arg[VEC];
            unsigned long long smask[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) { min1[e] = Num<T>::inf(); min2[e] = Num<T>::inf(); arg[e] = 0; smask[e] = 0; }
            for (int k = 0; k < deg; ++k)
            {
                const uint32_t col = Acc<SMEM, IdxT, 0>::ld(col0 + k * IS);
                const V o = VAcc<SMEM, T, 0>::ld(out_sub + col * RS);
                const V c = VAcc<SMEM, T, 0>::ld(c2v0 + k * CS);
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                {
                    const T v = o.e[e] - c.e[e];
                    par ^= (o.e[e] <= T(0)) ? (1u << e) : 0u;
                    smask[e] |= (unsigned long long)(Num<T>::hi(v) >> 31) << k;
                    const T a = Num<T>::abs(v);
                    const bool lt1 = a < min1[e], lt2 = a < min2[e];
                    min2[e] = lt1 ? min1[e] : (lt2 ? a : min2[e]);
                    arg[e] = lt1 ? k : arg[e];
                    min1[e] = lt1 ? a : min1[e];
                }
            }
            uint32_t tot[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) tot[e] = (uint32_t)__popcll(smask[e]) & 1u;
            for (int k = 0; k < deg; ++k)
            {
                V r;
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                {
                    const T mag = (k == arg[e]) ? min2[e] : min1[e];
                    const uint32_t s = tot[e] ^ (uint32_t)((smask[e] >> k) & 1ull);
                    r.e[e] = Num<T>::with_sign(mag, s << 31);
                }
                VAcc<SMEM, T, 0>::st(c2v0 + k * CS, r);
            }
        }
        else
        {
            // box-plus forward/backward with the forward values parked in the output slots: slot k first
            // receives F[k-1]; the backward sweep turns it into f(F[k-1], B[k+1]) (decoder.cpp:33-44)
            V Fp, B, vk;
            {
                const uint32_t col = Acc<SMEM, IdxT, 0>::ld(col0);
                const V o = VAcc<SMEM, T, 0>::ld(out_sub + col * RS);
                const V c = VAcc<SMEM, T, 0>::ld(c2v0);
#pragma unroll
                for (int e = 0; e < VEC; ++e) { Fp.e[e] = o.e[e] - c.e[e]; par ^= (o.e[e] <= T(0)) ? (1u << e) : 0u; }
            }
            // v[k] is needed again by the backward sweep: park it in the `B`-side by re-deriving it there
            // (out and the old c2v are gone by then), so keep v in a small local array instead
            V v[64];
            v[0] = Fp;
            for (int k = 1; k < deg; ++k)
            {
                const uint32_t col = Acc<SMEM, IdxT, 0>::ld(col0 + k * IS);
                const V o = VAcc<SMEM, T, 0>::ld(out_sub + col * RS);
                const V c = VAcc<SMEM, T, 0>::ld(c2v0 + k * CS);
#pragma unroll
                for (int e = 0; e < VEC; ++e) { vk.e[e] = o.e[e] - c.e[e]; par ^= (o.e[e] <= T(0)) ? (1u << e) : 0u; }
                v[k] = vk;
                VAcc<SMEM, T, 0>::st(c2v0 + k * CS, Fp); // F[k-1] (slot deg-1 thereby gets its final value)
#pragma unroll
                for (int e = 0; e < VEC; ++e) Fp.e[e] = boxplus(Fp.e[e], vk.e[e]);
            }
            B = v[deg - 1];
            for (int k = deg - 2; k >= 1; --k)
            {
                const V f = VAcc<SMEM, T, 0>::ld(c2v0 + k * CS);
                V r;
#pragma unroll
                for (int e = 0; e < VEC; ++e) { r.e[e] = boxplus(f.e[e], B.e[e]); B.e[e] = boxplus(B.e[e], v[k].e[e]); }
                VAcc<SMEM, T, 0>::st(c2v0 + k * CS, r);
            }
            VAcc<SMEM, T, 0>::st(c2v0, B);
        }
        return par;
    }

    // variable node: posterior = LLRin + sum of incoming c2v, strictly in file order (decoder.cpp:50-56)
    template <typename T, typename IdxT, bool SMEM, int LANES, int D>
    struct VnVec
    {
        typedef typename PtrOf<SMEM>::type P;
        typedef Vec<T> V;
        static constexpr int VEC = V::N, NPW = 32 / LANES, IS = NPW * (int)sizeof(IdxT), RS = 16 * LANES;
        template <int K> struct Step
        {
            static __device__ __forceinline__ void load(P c2v_sub, P slot0, V (&m)[D])
            {
                const uint32_t s = Acc<SMEM, IdxT, K * IS>::ld(slot0);
                m[K] = VAcc<SMEM, T, 0>::ld(c2v_sub + s * RS);
                if constexpr (K + 1 < D) Step<K + 1>::load(c2v_sub, slot0, m);
            }
        };
        static __device__ __forceinline__ V run(P c2v_sub, P slot0, V acc)
        {
            V m[D];
            Step<0>::load(c2v_sub, slot0, m);
#pragma unroll
            for (int k = 0; k < D; ++k)
            {
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc.e[e] += m[k].e[e];
            }
            return acc;
        }
    };

    template <typename T, typename IdxT, bool SMEM, int LANES>
    __device__ __forceinline__ Vec<T> vn_vec_any(typename PtrOf<SMEM>::type c2v_sub, typename PtrOf<SMEM>::type slot0, int deg, Vec<T> acc)
    {
        typedef Vec<T> V;
        constexpr int VEC = V::N, NPW = 32 / LANES, IS = NPW * (int)sizeof(IdxT), RS = 16 * LANES;
        int k = 0;
        for (; k + 4 <= deg; k += 4)
        {
            const uint32_t s0 = Acc<SMEM, IdxT, 0>::ld(slot0 + k * IS), s1 = Acc<SMEM, IdxT, IS>::ld(slot0 + k * IS);
            const uint32_t s2 = Acc<SMEM, IdxT, 2 * IS>::ld(slot0 + k * IS), s3 = Acc<SMEM, IdxT, 3 * IS>::ld(slot0 + k * IS);
            const V m0 = VAcc<SMEM, T, 0>::ld(c2v_sub + s0 * RS), m1 = VAcc<SMEM, T, 0>::ld(c2v_sub + s1 * RS);
            const V m2 = VAcc<SMEM, T, 0>::ld(c2v_sub + s2 * RS), m3 = VAcc<SMEM, T, 0>::ld(c2v_sub + s3 * RS);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { acc.e[e] += m0.e[e]; acc.e[e] += m1.e[e]; acc.e[e] += m2.e[e]; acc.e[e] += m3.e[e]; }
        }
        for (; k < deg; ++k)
        {
            const V m = VAcc<SMEM, T, 0>::ld(c2v_sub + (uint32_t)Acc<SMEM, IdxT, 0>::ld(slot0 + k * IS) * RS);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc.e[e] += m.e[e];
        }
        return acc;
    }

    // ------------------------------------------------------------------------------------------
    // the persistent kernel
    // ------------------------------------------------------------------------------------------
    template <typename T, typename IdxT, int ALG, bool SMEM, int LANES>
    __global__ void __launch_bounds__(ALG == ALG_MS ? 1024 : 512, 1) tile3_kernel(const K3Params p)
    {
        typedef typename PtrOf<SMEM>::type P;
        typedef Vec<T> V;
        constexpr int VEC = V::N, FPC = LANES * VEC, NPW = 32 / LANES;
        constexpr int TS = (int)sizeof(T), RS = 16 * LANES, IS1 = (int)sizeof(IdxT), IS = NPW * IS1;
        constexpr uint32_t ALL = (FPC == 32) ? 0xFFFFFFFFu : ((1u << FPC) - 1u), VMASK = (1u << VEC) - 1u;
        extern __shared__ __align__(16) unsigned char dyn_smem[];
        __shared__ unsigned long long s_frame[FPC], s_old[FPC];
        __shared__ unsigned long long s_cnt[5];
        __shared__ uint32_t s_err[2][FPC];
        __shared__ uint32_t s_synd[2];
        __shared__ uint2 s_ctrl[2]; // {frames at the iteration limit, frames with >= 1 completed iteration}
        __shared__ int s_ret[FPC];
        __shared__ uint32_t s_active, s_skip, s_next;

        const int tid = threadIdx.x, nthreads = blockDim.x;
        const int lane = tid & 31, warp = tid >> 5, warps = nthreads >> 5;
        const int sub = lane & (LANES - 1), j = lane / LANES;

        // ---- carve state and tables --------------------------------------------------------
        P c2v, out, llr, cn_task, vn_task, cn_col, vn_slot;
        if constexpr (SMEM)
        {
            uint32_t q = (uint32_t)__cvta_generic_to_shared(dyn_smem);
            const uint32_t a_c2v = q; q += RS * p.n_slots;
            const uint32_t a_out = q; q += RS * p.n_pos;
            const uint32_t a_llr = q; q += RS * p.n_pos;
            const uint32_t a_ct = q; q += 8 * p.cn_rounds * warps;
            const uint32_t a_vt = q; q += 8 * p.vn_rounds * warps;
            const uint32_t a_cc = q; q += (IS1 * p.n_slots + 15) & ~15;
            const uint32_t a_vs = q;
            for (int i = tid; i < 2 * p.cn_rounds * warps; i += nthreads) sts_u32<0>(a_ct + 4 * i, p.cn_task[i]);
            for (int i = tid; i < 2 * p.vn_rounds * warps; i += nthreads) sts_u32<0>(a_vt + 4 * i, p.vn_task[i]);
            for (int i = tid; i < p.n_slots; i += nthreads) Acc<true, IdxT, 0>::st(a_cc + IS1 * i, static_cast<const IdxT *>(p.cn_col)[i]);
            for (int i = tid; i < p.n_vslots; i += nthreads) Acc<true, IdxT, 0>::st(a_vs + IS1 * i, static_cast<const IdxT *>(p.vn_slot)[i]);
            c2v = a_c2v; out = a_out; llr = a_llr; cn_task = a_ct; vn_task = a_vt; cn_col = a_cc; vn_slot = a_vs;
        }
        else
        {
            unsigned char *q = p.state + p.state_stride * blockIdx.x;
            unsigned char *g_c2v = q; q += (size_t)RS * p.n_slots;
            unsigned char *g_out = q; q += (size_t)RS * p.n_pos;
            unsigned char *g_llr = q;
            c2v = g_c2v; out = g_out; llr = g_llr;
            cn_task = (unsigned char *)p.cn_task; vn_task = (unsigned char *)p.vn_task;
            cn_col = (unsigned char *)p.cn_col; vn_slot = (unsigned char *)p.vn_slot;
        }
        if (tid < 5) s_cnt[tid] = 0;
        if (tid < FPC) { s_err[0][tid] = 0; s_err[1][tid] = 0; s_frame[tid] = 0; s_old[tid] = 0; s_ret[tid] = 0; }
        if (tid == 0)
        {
            s_next = 0; s_active = 0; s_skip = 0; s_synd[0] = 0; s_synd[1] = 0;
            s_ctrl[0] = make_uint2(0, 0); s_ctrl[1] = make_uint2(0, 0);
        }
        __syncthreads();

        // per-frame iteration counter: lane g of warp 0 owns frame lane g
        int it = 0;
        uint32_t active = 0, skip = 0; // CTA-uniform copies of s_active / s_skip

        // Writes the decoder input of global frame gf into frame lane g (all threads of the CTA
        // cooperate), with the fresh-frame state: out = LLRin, c2v = +0.
        auto generate = [&](int g, unsigned long long gf)
        {
            const int eo = (g / VEC) * 16 + (g % VEC) * TS; // byte offset of lane g inside a record
            const P dl = llr + eo, dout = out + eo, dc = c2v + eo;
            for (int i = tid; i < p.n_slots; i += nthreads) Acc<SMEM, T, 0>::st(dc + i * RS, T(0));
            auto put = [&](int pos, T v)
            {
                Acc<SMEM, T, 0>::st(dl + pos * RS, v);
                Acc<SMEM, T, 0>::st(dout + pos * RS, v);
            };
            if (p.kind == SRC_LLR)
            {
                const double *src = p.llr_in + (size_t)gf * p.nc;
                for (int i = tid; i < p.nc; i += nthreads) put((int)p.var_pos[i], (T)src[i]);
                return;
            }
            const unsigned long long frame = p.frame0 + gf;
            if (p.kind == SRC_AWGN)
            { // y = sigma*z + 1 (all-zero codeword, BPSK +1), LLR = 2y/sigma^2 (src/sim/channel.cpp:62-68,88-92)
                const int npairs = (p.nct + 1) >> 1;
                for (int q = tid; q < npairs; q += nthreads)
                {
                    const u32x4 r = channel_block(p.seed, p.point, 0, frame, (uint32_t)q);
                    const double u1 = ((double)((((uint64_t)r.y << 32) | r.x) >> 11) + 1.0) * 0x1p-53;
                    const double u2 = (double)((((uint64_t)r.w << 32) | r.z) >> 11) * 0x1p-53;
                    const double rad = sqrt(-2.0 * log(u1));
                    double sn, cs;
                    sincos(6.283185307179586 * u2, &sn, &cs);
                    const double y0 = __dadd_rn(__dmul_rn(rad * cs, p.sigma), 1.0);
                    const double y1 = __dadd_rn(__dmul_rn(rad * sn, p.sigma), 1.0);
                    const int t = 2 * q;
                    put(p.tx_pos[t], (T)(__dmul_rn(2.0, y0) / p.sigma2));
                    if (t + 1 < p.nct) put(p.tx_pos[t + 1], (T)(__dmul_rn(2.0, y1) / p.sigma2));
                }
                for (int i = tid; i < p.n_punct; i += nthreads) put(p.punct_pos[i], T(0));
                for (int i = tid; i < p.n_short; i += nthreads) put(p.short_pos[i], (T)99999.9);
            }
            else
            { // BSC: y = x ^ Bernoulli(eps), LLR = delta*(1-2y) (src/sim/channel.cpp:123-162)
                const int nblk = (p.nct + 3) >> 2;
                for (int q = tid; q < nblk; q += nthreads)
                {
                    const u32x4 r = channel_block(p.seed, p.point, 0, frame, (uint32_t)q);
                    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        const int t = 4 * q + k;
                        if (t < p.nct) put(p.tx_pos[t], (T)((w[k] < p.thr) ? -p.delta : p.delta));
                    }
                }
                for (int i = tid; i < p.n_punct; i += nthreads) put(p.punct_pos[i], T(0));
                for (int i = tid; i < p.n_short; i += nthreads) put(p.short_pos[i], (T)p.delta);
            }
        };

        // Retires the frame lanes in `mask` and hands each a new frame if any is left.
        //   synd / started : syndrome flags and ">= 1 iteration done" flags valid for this decision
        //   err_par        : which s_err buffer holds the bit errors of the last variable phase
        //   as_skip        : the new frames must sit out the variable phase that follows
        auto retire_and_refill = [&](uint32_t mask, uint32_t synd, uint32_t started, int err_par, bool as_skip, bool first_fill)
        {
            if (warp == 0)
            {
                bool got = false;
                const bool mine = lane < FPC && ((mask >> lane) & 1u);
                if (mine && !first_fill)
                {
                    const bool conv = p.early_term && ((started >> lane) & 1u) && !((synd >> lane) & 1u);
                    const int ret = conv ? it - 1 : p.max_iter; // the reference breaks before ++I (decoder.cpp:66-77)
                    const uint32_t e = s_err[err_par][lane];
                    atomicAdd(&s_cnt[0], (unsigned long long)(e ? 1 : 0));
                    atomicAdd(&s_cnt[1], (unsigned long long)e);
                    atomicAdd(&s_cnt[2], 1ull);
                    atomicAdd(&s_cnt[3], (unsigned long long)ret);
                    atomicAdd(&s_cnt[4], (unsigned long long)it);
                    s_ret[lane] = ret;
                    s_old[lane] = s_frame[lane];
                }
                if (mine)
                {
                    const uint32_t k = s_next + (uint32_t)__popc(mask & ((1u << lane) - 1u));
                    const unsigned long long gf = (unsigned long long)blockIdx.x + (unsigned long long)gridDim.x * k;
                    got = gf < p.n_frames; // frame indices grow with k: the lanes that get one form a prefix of `mask`
                    if (got) s_frame[lane] = gf;
                    it = 0;
                }
                const uint32_t gm = __ballot_sync(0xffffffffu, got);
                if (lane == 0)
                {
                    s_next += (uint32_t)__popc(gm);
                    s_active = (active & ~mask) | gm;
                    s_skip = as_skip ? gm : 0u;
                }
            }
            __syncthreads();
            const uint32_t new_active = s_active;
            if (!first_fill && (p.llr_out || p.hard_out || p.iters_out))
            {
                for (int g = 0; g < FPC; ++g)
                    if ((mask >> g) & 1u)
                    {
                        const int eo = (g / VEC) * 16 + (g % VEC) * TS;
                        const size_t o = (size_t)s_old[g] * p.nc;
                        for (int i = tid; i < p.nc; i += nthreads)
                        {
                            const T v = Acc<SMEM, T, 0>::ld(out + eo + p.var_pos[i] * RS);
                            if (p.llr_out) p.llr_out[o + i] = (double)v;
                            if (p.hard_out) p.hard_out[o + i] = (v <= T(0)) ? 1 : 0; // decoder.cpp:58
                        }
                        if (tid == 0 && p.iters_out) p.iters_out[s_old[g]] = s_ret[g];
                    }
                __syncthreads();
            }
            for (int g = 0; g < FPC; ++g)
                if (((mask & new_active) >> g) & 1u) generate(g, s_frame[g]);
            __syncthreads();
            active = new_active;
            skip = s_skip;
        };

        retire_and_refill(ALL, 0, 0, 0, false, true);

        const P c2v_lane = c2v + lane * 16, c2v_sub = c2v + sub * 16, out_sub = out + sub * 16;
        const P out_lane = out + lane * 16, llr_lane = llr + lane * 16;
        const P cn_col_j = cn_col + j * IS1, vn_slot_j = vn_slot + j * IS1;

        for (uint32_t L = 0;; ++L)
        {
            const int par_i = (int)(L & 1u);
            if (!active) break; // CTA-uniform

            // ---- without early termination a frame at the iteration limit retires here, before a
            //      check phase is spent on it (its result is fixed: decoder.cpp:22,74-77)
            if (!p.early_term)
            {
                const uint32_t lim = s_ctrl[par_i].x & active;
                if (lim)
                {
                    retire_and_refill(lim, 0, 0, par_i ^ 1, false, false);
                    if (warp == 0 && lane == 0) s_ctrl[par_i].x &= ~lim;
                    if (!active) break;
                }
            }

            // ---- check-node phase (+ syndrome of the previous iteration's decisions) ----------
            uint32_t bad = 0;
            for (int r = 0; r < p.cn_rounds; ++r)
            {
                const uint2 t = TaskLd<SMEM>::ld(cn_task + 8 * (r * warps + warp));
                if (j >= (int)t.y) continue;
                const uint32_t p0 = t.x & 0xFFFFFFu;
                const int deg = (int)(t.x >> 24);
                const P c2v0 = c2v_lane + p0 * RS;
                const P col0 = cn_col_j + p0 * IS1;
                switch (deg) // warp-uniform
                {
                case 2: bad |= CnVec<T, IdxT, SMEM, LANES, ALG, 2>::run(out_sub, c2v0, col0); break;
                case 3: bad |= CnVec<T, IdxT, SMEM, LANES, ALG, 3>::run(out_sub, c2v0, col0); break;
                case 4: bad |= CnVec<T, IdxT, SMEM, LANES, ALG, 4>::run(out_sub, c2v0, col0); break;
                case 5: bad |= CnVec<T, IdxT, SMEM, LANES, ALG, 5>::run(out_sub, c2v0, col0); break;
                case 6: bad |= CnVec<T, IdxT, SMEM, LANES, ALG, 6>::run(out_sub, c2v0, col0); break;
                case 7: bad |= CnVec<T, IdxT, SMEM, LANES, ALG, 7>::run(out_sub, c2v0, col0); break;
                case 8: bad |= CnVec<T, IdxT, SMEM, LANES, ALG, 8>::run(out_sub, c2v0, col0); break;
                default: bad |= cn_vec_any<T, IdxT, SMEM, LANES, ALG>(out_sub, c2v0, col0, deg); break;
                }
            }
            // syndrome flags per frame lane: frame = sub*VEC + e
            {
                const uint32_t m = __reduce_or_sync(0xffffffffu, bad << (sub * VEC));
                if (lane == 0 && m) atomicOr(&s_synd[par_i], m);
            }
            __syncthreads(); // B

            // ---- decision: converged (decoder.cpp:66-72) or out of iterations ---------------------
            {
                const uint32_t synd = s_synd[par_i];
                const uint2 ctrl = s_ctrl[par_i];
                const uint32_t done = active & ((p.early_term ? (~synd & ctrl.y) : 0u) | ctrl.x);
                if (done) retire_and_refill(done, synd, ctrl.y, par_i ^ 1, true, false);
            }
            // bookkeeping for the variable phase that follows and the next decision
            const uint32_t live = active & ~skip;
            if (warp == 0)
            {
                if (lane < FPC && ((live >> lane) & 1u)) ++it;
                const bool act = lane < FPC && ((active >> lane) & 1u);
                const uint32_t started = __ballot_sync(0xffffffffu, act && it >= 1);
                const uint32_t limit = __ballot_sync(0xffffffffu, act && it >= p.max_iter);
                if (lane == 0) { s_ctrl[par_i ^ 1] = make_uint2(limit, started); s_synd[par_i ^ 1] = 0; }
                if (lane < FPC) s_err[par_i ^ 1][lane] = 0;
            }

            // ---- variable-node phase: posterior, hard decision, bit errors ----------------------
            const uint32_t mylive = (live >> (sub * VEC)) & VMASK;
            uint32_t err[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) err[e] = 0;
            if (live)
            {
                for (int r = 0; r < p.vn_rounds; ++r)
                {
                    const uint2 t = TaskLd<SMEM>::ld(vn_task + 8 * (r * warps + warp));
                    if (j >= (int)(t.y >> 24)) continue;
                    const uint32_t q0 = t.x & 0x7FFFFFu, vb = t.y & 0xFFFFFFu;
                    const int deg = (int)((t.x >> 23) & 0xFFu);
                    const P slot0 = vn_slot_j + q0 * IS1;
                    V acc = VAcc<SMEM, T, 0>::ld(llr_lane + vb * RS); // decoder.cpp:50
                    switch (deg)
                    {
                    case 0: break;
                    case 1: acc = VnVec<T, IdxT, SMEM, LANES, 1>::run(c2v_sub, slot0, acc); break;
                    case 2: acc = VnVec<T, IdxT, SMEM, LANES, 2>::run(c2v_sub, slot0, acc); break;
                    case 3: acc = VnVec<T, IdxT, SMEM, LANES, 3>::run(c2v_sub, slot0, acc); break;
                    case 4: acc = VnVec<T, IdxT, SMEM, LANES, 4>::run(c2v_sub, slot0, acc); break;
                    case 5: acc = VnVec<T, IdxT, SMEM, LANES, 5>::run(c2v_sub, slot0, acc); break;
                    case 6: acc = VnVec<T, IdxT, SMEM, LANES, 6>::run(c2v_sub, slot0, acc); break;
                    case 8: acc = VnVec<T, IdxT, SMEM, LANES, 8>::run(c2v_sub, slot0, acc); break;
                    default: acc = vn_vec_any<T, IdxT, SMEM, LANES>(c2v_sub, slot0, deg, acc); break;
                    }
                    const P o = out_lane + vb * RS;
                    if (mylive == VMASK) VAcc<SMEM, T, 0>::st(o, acc);
                    else
                    {
#pragma unroll
                        for (int e = 0; e < VEC; ++e)
                            if ((mylive >> e) & 1u) Acc<SMEM, T, 0>::st(o + e * TS, acc.e[e]);
                    }
                    if (t.x >> 31)
                    { // all-zero codeword: ldpcsim.cpp:184-188 (transmitted positions only)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) err[e] += (acc.e[e] <= T(0)) ? 1u : 0u;
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e)
            {
                uint32_t v = ((mylive >> e) & 1u) ? err[e] : 0u;
#pragma unroll
                for (int o = LANES; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane < LANES && v) atomicAdd(&s_err[par_i][lane * VEC + e], v);
            }
            skip = 0;
            __syncthreads(); // A: variable-phase writes visible to the next check phase
        }

        __syncthreads();
        if (tid < 5 && s_cnt[tid]) atomicAdd(&p.counters[tid], s_cnt[tid]);
    }
} // namespace b200
