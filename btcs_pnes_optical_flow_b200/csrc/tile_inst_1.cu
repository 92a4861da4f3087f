// Explicit instantiations of the tile-kernel launchers (farneback_fast.cuh) for half windows 3, 4, 5: split over several
// translation units so that the build compiles them in parallel.
#define BF_TILE_INSTANTIATE
#include "farneback_tile.cuh"

namespace bf {
BF_INSTANTIATE_TILE_MH(3)
BF_INSTANTIATE_TILE_MH(4)
BF_INSTANTIATE_TILE_MH(5)
}  // namespace bf
