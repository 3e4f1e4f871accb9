// sm_100a kernels of the decode hot path: Philox channel -> flooding min-sum / box-plus decode ->
// hard decision + error accounting, fused into ONE persistent kernel.
//
// Mapping ("tile"): a CTA owns FPC frame lanes; thread t serves frame lane f = t % FPC as node
// thread t / FPC.  All per-frame arrays are laid out [index][FPC] with the frame lane fastest, so a
// warp touching NPW = 32/FPC consecutive indices moves one contiguous 32*sizeof(T) block (conflict
// free in shared memory, fully coalesced in HBM).  State per lane: c2v message per edge slot, the
// posterior `out` per variable and the channel LLR per variable; v2c is never stored — it is
// recomputed as out - c2v, which is exactly the value the reference stores
// (src/decoding/decoder.cpp:60-63), so results stay bit-identical while one of the reference's two
// message arrays disappears.
//
// Node groups of equal degree are scheduled per warp (code.cpp), so the degree is warp-uniform and
// the node updates are dispatched to fully unrolled fixed-degree bodies: all loads of a node are
// issued back to back, message slots are reached with immediate offsets (slot stride = 32*sizeof(T)
// bytes), and no per-edge loop/address arithmetic remains.
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

    // ------------------------------------------------------------------------------------------
    // memory-space policies: shared window (32-bit byte addresses, explicit ld/st.shared so that no
    // generic-address arithmetic is rematerialised in the inner loops) or global memory
    // ------------------------------------------------------------------------------------------
    // (explicit overload set instead of partial specialisation: OFF must be an immediate)
#define B200_SMEM_LD(NAME, TYPE, PTXT, CONS)                                                        \
    template <int OFF> __device__ __forceinline__ TYPE NAME(uint32_t a)                             \
    {                                                                                               \
        TYPE v;                                                                                     \
        asm volatile("ld.shared." PTXT " %0, [%1+%2];" : "=" CONS(v) : "r"(a), "n"(OFF));           \
        return v;                                                                                   \
    }
#define B200_SMEM_ST(NAME, TYPE, PTXT, CONS)                                                        \
    template <int OFF> __device__ __forceinline__ void NAME(uint32_t a, TYPE v)                     \
    {                                                                                               \
        asm volatile("st.shared." PTXT " [%0+%1], %2;" ::"r"(a), "n"(OFF), CONS(v) : "memory");     \
    }
    B200_SMEM_LD(lds_f64, double, "f64", "d")
    B200_SMEM_LD(lds_f32, float, "f32", "f")
    B200_SMEM_LD(lds_u32, uint32_t, "u32", "r")
    B200_SMEM_ST(sts_f64, double, "f64", "d")
    B200_SMEM_ST(sts_f32, float, "f32", "f")
    B200_SMEM_ST(sts_u32, uint32_t, "u32", "r")
    template <int OFF> __device__ __forceinline__ uint32_t lds_u16(uint32_t a)
    {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(a), "n"(OFF));
        return v;
    }
    template <int OFF> __device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v)
    {
        asm volatile("st.shared.u16 [%0+%1], %2;" ::"r"(a), "n"(OFF), "h"((uint16_t)v) : "memory");
    }

    // typed front ends ----------------------------------------------------------------------------
    template <bool SMEM, typename V, int OFF> struct Acc;
    template <int OFF> struct Acc<true, double, OFF>
    {
        static __device__ __forceinline__ double ld(uint32_t a) { return lds_f64<OFF>(a); }
        static __device__ __forceinline__ void st(uint32_t a, double v) { sts_f64<OFF>(a, v); }
    };
    template <int OFF> struct Acc<true, float, OFF>
    {
        static __device__ __forceinline__ float ld(uint32_t a) { return lds_f32<OFF>(a); }
        static __device__ __forceinline__ void st(uint32_t a, float v) { sts_f32<OFF>(a, v); }
    };
    template <int OFF> struct Acc<true, uint32_t, OFF>
    {
        static __device__ __forceinline__ uint32_t ld(uint32_t a) { return lds_u32<OFF>(a); }
        static __device__ __forceinline__ void st(uint32_t a, uint32_t v) { sts_u32<OFF>(a, v); }
    };
    template <int OFF> struct Acc<true, uint16_t, OFF>
    {
        static __device__ __forceinline__ uint32_t ld(uint32_t a) { return lds_u16<OFF>(a); }
        static __device__ __forceinline__ void st(uint32_t a, uint32_t v) { sts_u16<OFF>(a, v); }
    };
    template <typename V, int OFF> struct Acc<false, V, OFF>
    {
        static __device__ __forceinline__ V ld(const unsigned char *a) { return *reinterpret_cast<const V *>(a + OFF); }
        static __device__ __forceinline__ void st(unsigned char *a, V v) { *reinterpret_cast<V *>(a + OFF) = v; }
    };
    template <bool SMEM> struct PtrOf { typedef uint32_t type; };
    template <> struct PtrOf<false> { typedef unsigned char *type; };

    template <typename T> struct Num;
    template <> struct Num<double>
    {
        static __device__ __forceinline__ uint32_t hi(double v) { return (uint32_t)__double2hiint(v); }
        static __device__ __forceinline__ double abs(double v) { return fabs(v); }
        // mag >= 0; sign bit taken from bit 31 of s
        static __device__ __forceinline__ double with_sign(double mag, uint32_t s) { return __hiloint2double(__double2hiint(mag) | (int)(s & 0x80000000u), __double2loint(mag)); }
        static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
        static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
        static __device__ __forceinline__ double exp_(double v) { return exp(v); }
        static __device__ __forceinline__ double log_(double v) { return log(v); }
    };
    template <> struct Num<float>
    {
        static __device__ __forceinline__ uint32_t hi(float v) { return __float_as_uint(v); }
        static __device__ __forceinline__ float abs(float v) { return fabsf(v); }
        static __device__ __forceinline__ float with_sign(float mag, uint32_t s) { return __uint_as_float(__float_as_uint(mag) | (s & 0x80000000u)); }
        static __device__ __forceinline__ float inf() { return __uint_as_float(0x7f800000u); }
        static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
        static __device__ __forceinline__ float exp_(float v) { return __expf(v); }
        static __device__ __forceinline__ float log_(float v) { return __logf(v); }
    };

    // pairwise box-plus with the Jacobian correction, decoder.h:12-15 (same expression, same order)
    template <typename T>
    __device__ __forceinline__ T boxplus(T x, T y)
    {
        const T ax = Num<T>::abs(x), ay = Num<T>::abs(y);
        const T m = (ay < ax) ? ay : ax;
        const T sm = Num<T>::with_sign(m, Num<T>::hi(x) ^ Num<T>::hi(y));
        const T num = T(1) + Num<T>::exp_(-Num<T>::abs(x + y));
        const T den = T(1) + Num<T>::exp_(-Num<T>::abs(x - y));
        return sm + Num<T>::log_(num / den);
    }

    // ------------------------------------------------------------------------------------------
    // check-node updates.  A node is described by
    //   src   : &src[0][f]                      (src = out, or llr on a frame's first iteration)
    //   c2v0  : &c2v[p0][f]   slot k at + k*CS  (CS = 32*sizeof(T): slots of a node are NPW apart)
    //   col0  : &cn_col[p0]   entry k at + k*IS (IS = NPW*sizeof(IdxT))
    // Both variants return the parity of the hard decisions of the check's variables (the syndrome
    // bit of the previous iteration's output, reference: decoder.h:47-64) — free with the gather.
    // ------------------------------------------------------------------------------------------
    template <typename T, typename IdxT, bool SMEM, int FPC, int ALG, int D>
    struct CnFixed
    {
        typedef typename PtrOf<SMEM>::type P;
        static constexpr int CS = 32 * (int)sizeof(T), IS = (32 / FPC) * (int)sizeof(IdxT), VS = FPC * (int)sizeof(T);

        template <int K> struct Step
        {
            static __device__ __forceinline__ void load(P src, P c2v0, P col0, bool first, T (&v)[D], uint32_t &par)
            {
                const uint32_t col = Acc<SMEM, IdxT, K * IS>::ld(col0);
                const T o = Acc<SMEM, T, 0>::ld(src + col * VS);
                T c = T(0);
                if (!first) c = Acc<SMEM, T, K * CS>::ld(c2v0);
                v[K] = o - c; // == the reference's stored v2c (decoder.cpp:62), or LLRin on the first pass (:18)
                par ^= (o <= T(0)) ? 1u : 0u;
                if constexpr (K + 1 < D) Step<K + 1>::load(src, c2v0, col0, first, v, par);
            }
            static __device__ __forceinline__ void store(P c2v0, const T (&r)[D])
            {
                Acc<SMEM, T, K * CS>::st(c2v0, r[K]);
                if constexpr (K + 1 < D) Step<K + 1>::store(c2v0, r);
            }
        };

        static __device__ __forceinline__ uint32_t run(P src, P c2v0, P col0, bool first)
        {
            T v[D], r[D];
            uint32_t par = 0;
            Step<0>::load(src, c2v0, col0, first, v, par);
            if (ALG == ALG_MS)
            {
                // min-sum: f = sign*sign*min (decoder.h:17-20) through the forward/backward recursion of
                // decoder.cpp:30-44.  Magnitudes: exact prefix/suffix minima; signs: XOR of sign BITS
                // (std::signbit semantics, -0.0 is negative).
                uint32_t sx = 0;
                T a[D], pre[D], suf[D];
#pragma unroll
                for (int k = 0; k < D; ++k) { sx ^= Num<T>::hi(v[k]); a[k] = Num<T>::abs(v[k]); }
                pre[0] = a[0];
#pragma unroll
                for (int k = 1; k < D; ++k) pre[k] = Num<T>::min_(pre[k - 1], a[k]);
                suf[D - 1] = a[D - 1];
#pragma unroll
                for (int k = D - 2; k >= 0; --k) suf[k] = Num<T>::min_(suf[k + 1], a[k]);
#pragma unroll
                for (int k = 0; k < D; ++k)
                {
                    const T mag = (k == 0) ? suf[1] : (k == D - 1) ? pre[D - 2] : Num<T>::min_(pre[k - 1], suf[k + 1]);
                    r[k] = Num<T>::with_sign(mag, sx ^ Num<T>::hi(v[k]));
                }
            }
            else
            {
                // sum-product: the reference's forward/backward box-plus recursion, file order
                T F[D];
                F[0] = v[0];
#pragma unroll
                for (int k = 1; k < D; ++k) F[k] = boxplus(F[k - 1], v[k]);
                T B = v[D - 1];
                r[D - 1] = F[D - 2];
#pragma unroll
                for (int k = D - 2; k >= 1; --k) { r[k] = boxplus(F[k - 1], B); B = boxplus(B, v[k]); }
                r[0] = B;
            }
            Step<0>::store(c2v0, r);
            return par;
        }
    };

    // arbitrary degree (<= 64): running min1/min2 + sign mask for min-sum, parked forward values for box-plus
    template <typename T, typename IdxT, bool SMEM, int FPC, int ALG>
    __device__ __noinline__ uint32_t cn_any(typename PtrOf<SMEM>::type src, typename PtrOf<SMEM>::type c2v0, typename PtrOf<SMEM>::type col0,
                                            int deg, bool first)
    {
        constexpr int CS = 32 * (int)sizeof(T), IS = (32 / FPC) * (int)sizeof(IdxT), VS = FPC * (int)sizeof(T);
        uint32_t par = 0;
        if (ALG == ALG_MS)
        {
            T min1 = Num<T>::inf(), min2 = Num<T>::inf();
            int arg = 0;
            unsigned long long smask = 0;
            for (int k = 0; k < deg; ++k)
            {
                const uint32_t col = Acc<SMEM, IdxT, 0>::ld(col0 + k * IS);
                const T o = Acc<SMEM, T, 0>::ld(src + col * VS);
                const T c = first ? T(0) : Acc<SMEM, T, 0>::ld(c2v0 + k * CS);
                const T v = o - c;
                par ^= (o <= T(0)) ? 1u : 0u;
                smask |= (unsigned long long)(Num<T>::hi(v) >> 31) << k;
                const T a = Num<T>::abs(v);
                const bool lt1 = a < min1, lt2 = a < min2;
                min2 = lt1 ? min1 : (lt2 ? a : min2);
                arg = lt1 ? k : arg;
                min1 = lt1 ? a : min1;
            }
            const uint32_t tot = (uint32_t)__popcll(smask) & 1u;
            for (int k = 0; k < deg; ++k)
            {
                const T mag = (k == arg) ? min2 : min1;
                const uint32_t s = tot ^ (uint32_t)((smask >> k) & 1ull);
                Acc<SMEM, T, 0>::st(c2v0 + k * CS, Num<T>::with_sign(mag, s << 31));
            }
        }
        else
        {
            T v[64];
            for (int k = 0; k < deg; ++k)
            {
                const uint32_t col = Acc<SMEM, IdxT, 0>::ld(col0 + k * IS);
                const T o = Acc<SMEM, T, 0>::ld(src + col * VS);
                const T c = first ? T(0) : Acc<SMEM, T, 0>::ld(c2v0 + k * CS);
                v[k] = o - c;
                par ^= (o <= T(0)) ? 1u : 0u;
            }
            T Fp = v[0]; // F[k-1] while visiting k
            for (int k = 1; k < deg; ++k)
            {
                Acc<SMEM, T, 0>::st(c2v0 + k * CS, Fp); // park F[k-1] in slot k (slot deg-1 thereby gets its final value)
                Fp = boxplus(Fp, v[k]);
            }
            T B = v[deg - 1];
            for (int k = deg - 2; k >= 1; --k)
            {
                const T f = Acc<SMEM, T, 0>::ld(c2v0 + k * CS);
                Acc<SMEM, T, 0>::st(c2v0 + k * CS, boxplus(f, B));
                B = boxplus(B, v[k]);
            }
            Acc<SMEM, T, 0>::st(c2v0, B);
        }
        return par;
    }

    // variable node: posterior = LLRin + sum of incoming c2v, strictly in file order (decoder.cpp:50-56)
    template <typename T, typename IdxT, bool SMEM, int FPC, int D>
    struct VnFixed
    {
        typedef typename PtrOf<SMEM>::type P;
        static constexpr int IS = (32 / FPC) * (int)sizeof(IdxT), VS = FPC * (int)sizeof(T);
        template <int K> struct Step
        {
            static __device__ __forceinline__ void load(P c2v, P slot0, T (&m)[D])
            {
                const uint32_t s = Acc<SMEM, IdxT, K * IS>::ld(slot0);
                m[K] = Acc<SMEM, T, 0>::ld(c2v + s * VS);
                if constexpr (K + 1 < D) Step<K + 1>::load(c2v, slot0, m);
            }
        };
        static __device__ __forceinline__ T run(P c2v, P slot0, T acc)
        {
            T m[D];
            Step<0>::load(c2v, slot0, m);
#pragma unroll
            for (int k = 0; k < D; ++k) acc += m[k];
            return acc;
        }
    };

    template <typename T, typename IdxT, bool SMEM, int FPC>
    __device__ __forceinline__ T vn_any(typename PtrOf<SMEM>::type c2v, typename PtrOf<SMEM>::type slot0, int deg, T acc)
    {
        constexpr int IS = (32 / FPC) * (int)sizeof(IdxT), VS = FPC * (int)sizeof(T);
        int k = 0;
        for (; k + 4 <= deg; k += 4)
        {
            const uint32_t s0 = Acc<SMEM, IdxT, 0>::ld(slot0 + k * IS), s1 = Acc<SMEM, IdxT, IS>::ld(slot0 + k * IS);
            const uint32_t s2 = Acc<SMEM, IdxT, 2 * IS>::ld(slot0 + k * IS), s3 = Acc<SMEM, IdxT, 3 * IS>::ld(slot0 + k * IS);
            const T m0 = Acc<SMEM, T, 0>::ld(c2v + s0 * VS), m1 = Acc<SMEM, T, 0>::ld(c2v + s1 * VS);
            const T m2 = Acc<SMEM, T, 0>::ld(c2v + s2 * VS), m3 = Acc<SMEM, T, 0>::ld(c2v + s3 * VS);
            acc += m0; acc += m1; acc += m2; acc += m3;
        }
        for (; k < deg; ++k) acc += Acc<SMEM, T, 0>::ld(c2v + (uint32_t)Acc<SMEM, IdxT, 0>::ld(slot0 + k * IS) * VS);
        return acc;
    }

    // ------------------------------------------------------------------------------------------
    // the persistent tile kernel
    // ------------------------------------------------------------------------------------------
    template <typename T, typename IdxT, int ALG, bool SMEM, int FPC>
    __global__ void __launch_bounds__(ALG == ALG_MS ? 1024 : 512, 1) tile_kernel(const KParams p)
    {
        typedef typename PtrOf<SMEM>::type P;
        constexpr int FSHIFT = (FPC == 32) ? 5 : (FPC == 16) ? 4 : (FPC == 8) ? 3 : (FPC == 4) ? 2 : (FPC == 2) ? 1 : 0;
        constexpr int NPW = 32 / FPC;
        constexpr int TS = (int)sizeof(T), VS = FPC * TS, IS1 = (int)sizeof(IdxT);
        extern __shared__ __align__(16) unsigned char dyn_smem[];
        __shared__ unsigned long long s_frame[32];
        __shared__ unsigned long long s_cnt[5];
        __shared__ uint32_t s_synd[32], s_err[32], s_newstate[32];
        __shared__ int s_ret[32];
        __shared__ uint32_t s_done_mask, s_next;

        const int tid = threadIdx.x, nthreads = blockDim.x;
        const int f = tid & (FPC - 1);
        const int nth = tid >> FSHIFT;     // node thread
        const int NT = nthreads >> FSHIFT; // node threads per CTA
        const int lane = tid & 31;
        const int nc = p.nc;

        // ---- carve state and tables --------------------------------------------------------
        P c2v, out, llr, cn_desc, vn_desc, cn_col, vn_slot, vn_id;
        if constexpr (SMEM)
        {
            uint32_t q = (uint32_t)__cvta_generic_to_shared(dyn_smem);
            uint32_t a_c2v = q; q += TS * p.n_slots * FPC;
            uint32_t a_out = q; q += TS * nc * FPC;
            uint32_t a_llr = q; q += TS * nc * FPC;
            uint32_t a_cd = q; q += 4 * p.cn_rounds * NT;
            uint32_t a_vd = q; q += 4 * p.vn_rounds * NT;
            uint32_t a_cc = q; q += IS1 * p.n_slots;
            uint32_t a_vs = q; q += IS1 * p.n_vslots;
            uint32_t a_vi = q;
            for (int i = tid; i < p.cn_rounds * NT; i += nthreads) sts_u32<0>(a_cd + 4 * i, p.cn_desc[i]);
            for (int i = tid; i < p.vn_rounds * NT; i += nthreads)
            {
                sts_u32<0>(a_vd + 4 * i, p.vn_desc[i]);
                Acc<true, IdxT, 0>::st(a_vi + IS1 * i, static_cast<const IdxT *>(p.vn_id)[i]);
            }
            for (int i = tid; i < p.n_slots; i += nthreads) Acc<true, IdxT, 0>::st(a_cc + IS1 * i, static_cast<const IdxT *>(p.cn_col)[i]);
            for (int i = tid; i < p.n_vslots; i += nthreads) Acc<true, IdxT, 0>::st(a_vs + IS1 * i, static_cast<const IdxT *>(p.vn_slot)[i]);
            c2v = a_c2v; out = a_out; llr = a_llr; cn_desc = a_cd; vn_desc = a_vd; cn_col = a_cc; vn_slot = a_vs; vn_id = a_vi;
        }
        else
        {
            unsigned char *q = p.state + p.state_stride * blockIdx.x;
            unsigned char *g_c2v = q; q += (size_t)TS * p.n_slots * FPC;
            unsigned char *g_out = q; q += (size_t)TS * nc * FPC;
            unsigned char *g_llr = q;
            c2v = g_c2v; out = g_out; llr = g_llr;
            cn_desc = (unsigned char *)p.cn_desc; vn_desc = (unsigned char *)p.vn_desc;
            cn_col = (unsigned char *)p.cn_col; vn_slot = (unsigned char *)p.vn_slot; vn_id = (unsigned char *)p.vn_id;
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
            const P dst = llr + g * TS;
            if (p.kind == SRC_LLR)
            {
                const double *src = p.llr_in + (size_t)gf * nc;
                for (int i = tid; i < nc; i += nthreads) Acc<SMEM, T, 0>::st(dst + i * VS, (T)src[i]);
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
                    Acc<SMEM, T, 0>::st(dst + p.bit_pos[t] * VS, (T)(__dmul_rn(2.0, y0) / p.sigma2));
                    if (t + 1 < p.nct) Acc<SMEM, T, 0>::st(dst + p.bit_pos[t + 1] * VS, (T)(__dmul_rn(2.0, y1) / p.sigma2));
                }
                for (int i = tid; i < p.n_punct; i += nthreads) Acc<SMEM, T, 0>::st(dst + p.punct[i] * VS, T(0));
                for (int i = tid; i < p.n_short; i += nthreads) Acc<SMEM, T, 0>::st(dst + p.shorten[i] * VS, (T)99999.9);
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
                        if (t < p.nct) Acc<SMEM, T, 0>::st(dst + p.bit_pos[t] * VS, (T)((w[q] < p.thr) ? -p.delta : p.delta));
                    }
                }
                for (int i = tid; i < p.n_punct; i += nthreads) Acc<SMEM, T, 0>::st(dst + p.punct[i] * VS, T(0));
                for (int i = tid; i < p.n_short; i += nthreads) Acc<SMEM, T, 0>::st(dst + p.shorten[i] * VS, (T)p.delta);
            }
        };

        // Retires the lanes in `mask` (results + accounting were recorded by the caller in s_ret /
        // s_cnt), hands each a new frame if any is left and regenerates its input.
        auto retire_and_refill = [&](uint32_t mask, uint32_t new_state, bool write_outputs)
        {
            if (write_outputs && (p.llr_out || p.hard_out || p.iters_out))
            {
                for (int g = 0; g < FPC; ++g)
                    if ((mask >> g) & 1u)
                    {
                        const size_t o = (size_t)s_frame[g] * nc;
                        for (int i = tid; i < nc; i += nthreads)
                        {
                            const T v = Acc<SMEM, T, 0>::ld(out + i * VS + g * TS);
                            if (p.llr_out) p.llr_out[o + i] = (double)v;
                            if (p.hard_out) p.hard_out[o + i] = (v <= T(0)) ? 1 : 0; // decoder.cpp:58
                        }
                        if (tid == 0 && p.iters_out) p.iters_out[s_frame[g]] = s_ret[g];
                    }
                __syncthreads();
            }
            if (tid == 0)
            {
                for (int g = 0; g < FPC; ++g)
                    if ((mask >> g) & 1u)
                    {
                        const unsigned long long gf = (unsigned long long)blockIdx.x + (unsigned long long)gridDim.x * s_next;
                        if (gf < p.n_frames) { s_frame[g] = gf; s_newstate[g] = new_state; ++s_next; }
                        else s_newstate[g] = 0;
                    }
            }
            __syncthreads();
            for (int g = 0; g < FPC; ++g)
                if (((mask >> g) & 1u) && s_newstate[g]) generate(g, s_frame[g]);
            if ((mask >> f) & 1u) { st = s_newstate[f]; it = 0; }
        };

        retire_and_refill(FPC == 32 ? 0xFFFFFFFFu : ((1u << FPC) - 1u), 1u, false);

        // bit pattern of the warp lanes that serve frame lane 0
        uint32_t lane_pattern = 0;
#pragma unroll
        for (int b = 0; b < 32; b += FPC) lane_pattern |= 1u << b;

        const P c2v_f = c2v + f * TS, out_f = out + f * TS, llr_f = llr + f * TS;

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
                    if (tid < FPC && st == 1 && it >= p.max_iter)
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
                const P src = first ? llr_f : out_f;
                for (int r = 0; r < p.cn_rounds; ++r)
                {
                    const uint32_t d = Acc<SMEM, uint32_t, 0>::ld(cn_desc + 4 * (r * NT + nth));
                    if (d == IDLE) continue;
                    const uint32_t p0 = d & 0xFFFFFFu;
                    const int deg = (int)(d >> 24);
                    const P c2v0 = c2v_f + p0 * VS;
                    const P col0 = cn_col + p0 * IS1;
                    switch (deg) // warp-uniform: a warp's node group has one degree
                    {
                    case 2: par |= CnFixed<T, IdxT, SMEM, FPC, ALG, 2>::run(src, c2v0, col0, first); break;
                    case 3: par |= CnFixed<T, IdxT, SMEM, FPC, ALG, 3>::run(src, c2v0, col0, first); break;
                    case 4: par |= CnFixed<T, IdxT, SMEM, FPC, ALG, 4>::run(src, c2v0, col0, first); break;
                    case 5: par |= CnFixed<T, IdxT, SMEM, FPC, ALG, 5>::run(src, c2v0, col0, first); break;
                    case 6: par |= CnFixed<T, IdxT, SMEM, FPC, ALG, 6>::run(src, c2v0, col0, first); break;
                    case 7: par |= CnFixed<T, IdxT, SMEM, FPC, ALG, 7>::run(src, c2v0, col0, first); break;
                    case 8: par |= CnFixed<T, IdxT, SMEM, FPC, ALG, 8>::run(src, c2v0, col0, first); break;
                    default: par |= cn_any<T, IdxT, SMEM, FPC, ALG>(src, c2v0, col0, deg, first); break;
                    }
                }
            }
            // warp-ballot syndrome test: one vote per thread, folded per frame lane
            {
                const uint32_t b = __ballot_sync(0xffffffffu, par != 0);
                if (lane < FPC && (b & (lane_pattern << lane))) s_synd[lane] = 1;
            }
            __syncthreads();

            // ---- decision: converged (reference: decoder.cpp:66-72) or out of iterations --------
            if (tid < 32)
            {
                bool fin = false;
                if (tid < FPC)
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
                    const uint32_t d = Acc<SMEM, uint32_t, 0>::ld(vn_desc + 4 * (r * NT + nth));
                    if (d == IDLE) continue;
                    const uint32_t id = Acc<SMEM, IdxT, 0>::ld(vn_id + IS1 * (r * NT + nth));
                    const uint32_t q0 = d & 0x7FFFFFu;
                    const int deg = (int)((d >> 23) & 0xFFu);
                    const P slot0 = vn_slot + q0 * IS1;
                    T acc = Acc<SMEM, T, 0>::ld(llr_f + id * VS); // decoder.cpp:50
                    switch (deg)
                    {
                    case 0: break;
                    case 1: acc = VnFixed<T, IdxT, SMEM, FPC, 1>::run(c2v_f, slot0, acc); break;
                    case 2: acc = VnFixed<T, IdxT, SMEM, FPC, 2>::run(c2v_f, slot0, acc); break;
                    case 3: acc = VnFixed<T, IdxT, SMEM, FPC, 3>::run(c2v_f, slot0, acc); break;
                    case 4: acc = VnFixed<T, IdxT, SMEM, FPC, 4>::run(c2v_f, slot0, acc); break;
                    case 5: acc = VnFixed<T, IdxT, SMEM, FPC, 5>::run(c2v_f, slot0, acc); break;
                    case 6: acc = VnFixed<T, IdxT, SMEM, FPC, 6>::run(c2v_f, slot0, acc); break;
                    case 8: acc = VnFixed<T, IdxT, SMEM, FPC, 8>::run(c2v_f, slot0, acc); break;
                    case 15: acc = VnFixed<T, IdxT, SMEM, FPC, 15>::run(c2v_f, slot0, acc); break;
                    default: acc = vn_any<T, IdxT, SMEM, FPC>(c2v_f, slot0, deg, acc); break;
                    }
                    Acc<SMEM, T, 0>::st(out_f + id * VS, acc);
                    err += ((d >> 31) && acc <= T(0)) ? 1u : 0u; // all-zero codeword: ldpcsim.cpp:184-188
                }
                ++it;
            }
            else if (st == 2) st = 1;
#pragma unroll
            for (int o = FPC; o < 32; o <<= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
            if (lane < FPC && err) atomicAdd(&s_err[lane], err);
        }

        __syncthreads();
        if (tid < 5 && s_cnt[tid]) atomicAdd(&p.counters[tid], s_cnt[tid]);
    }
} // namespace b200
