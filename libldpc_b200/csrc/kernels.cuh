// Device-side building blocks shared by the sm_100a kernels of the decode hot path (tile4.cuh: BP / min-sum,
// bec_kernel.cuh: erasure decoder, engine.cu: stand-alone channel kernel): the counter-based Philox4x32-10
// channel stream, memory-space access policies (shared window with explicit ld/st.shared, or global memory),
// numeric helpers and the pairwise box-plus of the reference (src/decoding/decoder.h:12-15).
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

} // namespace b200
