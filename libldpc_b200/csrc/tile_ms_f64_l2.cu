// tile kernel instantiation: T = double, algorithm = ALG_MS, lanes per node = 2
#include "tile_launch.cuh"
namespace b200
{
    B200_DEFINE_TILE_LANES(double, ALG_MS, 2)
}
