// Probe: Tensor Memory as per-thread private scratch (tcgen05.st / tcgen05.ld, shape 32x32b.x4) next to
// shared-memory traffic.  Checks data integrity for 32 warps sharing the 4 lane partitions by column
// ranges, and times (a) LDS.128-only streaming, (b) LDTM-only, (c) both interleaved.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int MODE> // 0: smem only, 1: tmem only, 2: both
__global__ void __launch_bounds__(1024, 1) probe(int iters, uint32_t *out, unsigned long long *cycles, int *errors)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_base;
    const int nw = blockDim.x >> 5, per = 512 / ((nw + 3) / 4); // columns per warp
    const uint32_t my = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * per;
    // integrity: every thread writes a unique pattern to all of its columns, reads it back after a barrier
    for (int c = 0; c < per; c += 4) tm_st4(my + c, tid * 1000 + c, tid * 1000 + c + 1, tid * 1000 + c + 2, tid * 1000 + c + 3);
    tm_wait_st();
    __syncthreads();
    int bad = 0;
    for (int c = 0; c < per; c += 4)
    {
        uint32_t a, b, cc, d;
        tm_ld4(my + c, a, b, cc, d);
        tm_wait_ld();
        bad += (a != tid * 1000 + c) + (b != tid * 1000 + c + 1) + (cc != tid * 1000 + c + 2) + (d != tid * 1000 + c + 3);
    }
    if (bad) atomicAdd(errors, bad);
    // throughput
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem) + tid * 16;
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
            uint32_t a = 0, b = 0, c = 0, d = 0, e = 0, f = 0, g = 0, h = 0;
            if (MODE != 1) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(sa + ((i * 4 + u) & 7) * 16384));
            if (MODE != 0) { tm_ld4(my + ((i * 4 + u) * 4) % per, e, f, g, h); tm_wait_ld(); }
            acc += a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
        }
    }
    const unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

int main()
{
    uint32_t *out; unsigned long long *cyc; int *err;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
    const int iters = 20000, smem = 8 * 16384 + 16384;
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 3; ++mode)
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (mode == 0) probe<0><<<148, 1024, smem>>>(iters, out, cyc, err);
        if (mode == 1) probe<1><<<148, 1024, smem>>>(iters, out, cyc, err);
        if (mode == 2) probe<2><<<148, 1024, smem>>>(iters, out, cyc, err);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long c0; int herr;
        cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
        const double bytes = 148.0 * 1024 * 16 * 4 * iters;
        printf("mode %d (%s): %s, %.3f ms, %llu cycles, integrity errors %d, %.1f B/clk/SM per stream, %.2f TB/s per stream\n", mode,
               mode == 0 ? "LDS.128 only" : mode == 1 ? "LDTM.x4 only" : "LDS.128 + LDTM.x4", cudaGetErrorString(e), ms, c0, herr,
               1024.0 * 16 * 4 * iters / (double)c0, bytes / (ms * 1e-3) / 1e12);
    }
    return 0;
}
