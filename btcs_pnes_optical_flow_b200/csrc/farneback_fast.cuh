// Specialised (compile-time window / poly_n) kernels for the configurations BASELINE.json names.
// Filled in after the generic path is parity-green; the dispatch predicates below gate them.
#pragma once
#include "farneback_kernels.cuh"

namespace bf {

inline bool polyexp_fast_supported(int /*n*/, int /*pitch*/) { return false; }
inline void launch_polyexp_fast(const float*, int, size_t, int, int, float*, size_t, size_t, int, int, int,
                                const PolyCoef&, cudaStream_t) {}

inline bool blur_solve_fast_supported(const WinCoef&, int /*pitch*/) { return false; }
inline int blur_solve_fast_ncta(int, int) { return 0; }
inline void launch_blur_solve_fast(const BlurSolveArgs&, const WinCoef&, int, cudaStream_t) {}

}  // namespace bf
