// Explicit instantiations of the tile-kernel launchers (farneback_fast.cuh) for half windows 14, 15, 16: split over several
// translation units so that the build compiles them in parallel.
#define BF_TILE_INSTANTIATE
#include "farneback_tile.cuh"

namespace bf {
BF_INSTANTIATE_TILE_MH(14)
BF_INSTANTIATE_TILE_MH(15)
BF_INSTANTIATE_TILE_MH(16)
}  // namespace bf
