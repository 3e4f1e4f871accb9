// FP64 pipe micro-benchmark: dependent DFMA latency and throughput as a function of independent chains per warp and warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/micro/dfma_probe profiles/micro/dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters, double a, double b, long long *cyc)
{
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = __fma_rn(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH> void run(int threads, double *d, long long *dc)
{
    const int iters = 2000;
    k<CH><<<148, threads>>>(d, iters, 0.999, 1e-3, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 8.0 * CH); // cycles per DFMA per warp
    const int wps = threads / 32 / 4 > 0 ? threads / 32 / 4 : 1;
    printf("chains %d warps/SM %2d: %.2f cycles per warp-DFMA, %.2f cycles per DFMA per SMSP (%d warps per SMSP)\n", CH, threads / 32, per, per / wps, wps);
}
int main()
{
    double *d; long long *dc;
    cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    for (int threads : {32, 128, 256, 512, 1024})
    {
        run<1>(threads, d, dc); run<2>(threads, d, dc); run<3>(threads, d, dc); run<4>(threads, d, dc); run<6>(threads, d, dc); run<8>(threads, d, dc);
    }
    return 0;
}
