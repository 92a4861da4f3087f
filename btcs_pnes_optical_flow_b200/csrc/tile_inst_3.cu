// Explicit instantiations of the tile-kernel launchers (farneback_fast.cuh) for half windows 11, 12, 13: split over several
// translation units so that the build compiles them in parallel.
#define BF_TILE_INSTANTIATE
#include "farneback_tile.cuh"

namespace bf {
BF_INSTANTIATE_TILE_MH(11)
BF_INSTANTIATE_TILE_MH(12)
BF_INSTANTIATE_TILE_MH(13)
}  // namespace bf
