// Explicit instantiations of the tile-kernel launchers (farneback_fast.cuh) for half windows 6, 8, 9: split over several
// translation units so that the build compiles them in parallel.
#define BF_TILE_INSTANTIATE
#include "farneback_tile.cuh"

namespace bf {
BF_INSTANTIATE_TILE_MH(6)
BF_INSTANTIATE_TILE_MH(8)
BF_INSTANTIATE_TILE_MH(9)
}  // namespace bf
