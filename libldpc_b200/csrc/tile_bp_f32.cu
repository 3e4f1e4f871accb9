// tile kernel family instantiation: T = float, algorithm = ALG_BP
#include "tile_launch.cuh"
namespace b200
{
    B200_DEFINE_TILE_FAMILY(float, ALG_BP)
}
