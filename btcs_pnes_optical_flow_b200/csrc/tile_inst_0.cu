// Explicit instantiations of the tile-kernel launchers (farneback_fast.cuh) for half windows 7, 10, 2: split over several
// translation units so that the build compiles them in parallel.
#define BF_TILE_INSTANTIATE
#include "farneback_tile.cuh"

namespace bf {
BF_INSTANTIATE_TILE_MH(7)
BF_INSTANTIATE_TILE_MH(10)
BF_INSTANTIATE_TILE_MH(2)
}  // namespace bf
