// tile kernel instantiation: T = double, algorithm = ALG_BP, lanes per node = 4
#include "tile_launch.cuh"
namespace b200
{
    B200_DEFINE_TILE_LANES(double, ALG_BP, 4)
}
