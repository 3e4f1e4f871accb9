// Launch front end of the tile kernel family.  Each (T, ALG) pair is instantiated in its own
// translation unit (tile_*.cu) so that the build can compile them in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

#include "kernels.cuh"

namespace b200
{
    // SMEM residency uses 16-bit indices, global residency 32-bit indices.
    template <typename T, int ALG>
    void launch_tile_family(const KParams &kp, bool smem, int fpc, int ctas, int threads, size_t smem_bytes, cudaStream_t s);

    template <typename T, typename IdxT, int ALG, bool SMEM, int FPC>
    void launch_tile_one(const KParams &kp, int ctas, int threads, size_t smem_bytes, cudaStream_t s)
    {
        static bool attr_set = false;
        if (SMEM && !attr_set)
        {
            cudaError_t e = cudaFuncSetAttribute(tile_kernel<T, IdxT, ALG, SMEM, FPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 2048);
            if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e));
            attr_set = true;
        }
        tile_kernel<T, IdxT, ALG, SMEM, FPC><<<ctas, threads, SMEM ? smem_bytes : 0, s>>>(kp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " (tile kernel launch)");
    }

#define B200_DEFINE_TILE_FAMILY(T, ALG)                                                                                     \
    template <>                                                                                                             \
    void launch_tile_family<T, ALG>(const KParams &kp, bool smem, int fpc, int ctas, int threads, size_t smem_bytes, cudaStream_t s) \
    {                                                                                                                       \
        if (smem)                                                                                                           \
        {                                                                                                                   \
            switch (fpc)                                                                                                    \
            {                                                                                                               \
            case 4: launch_tile_one<T, uint16_t, ALG, true, 4>(kp, ctas, threads, smem_bytes, s); return;                   \
            case 8: launch_tile_one<T, uint16_t, ALG, true, 8>(kp, ctas, threads, smem_bytes, s); return;                   \
            case 16: launch_tile_one<T, uint16_t, ALG, true, 16>(kp, ctas, threads, smem_bytes, s); return;                 \
            case 32: launch_tile_one<T, uint16_t, ALG, true, 32>(kp, ctas, threads, smem_bytes, s); return;                 \
            }                                                                                                               \
        }                                                                                                                   \
        else                                                                                                                \
        {                                                                                                                   \
            switch (fpc)                                                                                                    \
            {                                                                                                               \
            case 4: launch_tile_one<T, uint32_t, ALG, false, 4>(kp, ctas, threads, 0, s); return;                           \
            case 8: launch_tile_one<T, uint32_t, ALG, false, 8>(kp, ctas, threads, 0, s); return;                           \
            case 16: launch_tile_one<T, uint32_t, ALG, false, 16>(kp, ctas, threads, 0, s); return;                         \
            case 32: launch_tile_one<T, uint32_t, ALG, false, 32>(kp, ctas, threads, 0, s); return;                         \
            }                                                                                                               \
        }                                                                                                                   \
        throw std::runtime_error("frames_per_cta must be 4, 8, 16 or 32");                                                  \
    }
} // namespace b200
