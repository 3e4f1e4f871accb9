// tile kernel family instantiation: T = double, algorithm = ALG_BP
#include "tile_launch.cuh"
namespace b200
{
    B200_DEFINE_TILE_FAMILY(double, ALG_BP)
}
