// tile kernel instantiation: T = float, algorithm = ALG_MS, lanes per node = 1
#include "tile_launch.cuh"
namespace b200
{
    B200_DEFINE_TILE_LANES(float, ALG_MS, 1)
}
