// NaN-robust zero-phase band-pass on the device (replaces bandpass_nanrobust, /root/reference/optical_PCA.py:96-121, i.e.
// scipy.signal.sosfiltfilt per contiguous finite run).  float64 like the reference.
//
// One warp per series.  The warp scans the series for finite runs (ballot); each run of >= min_len samples is filtered
// exactly like scipy: odd extension by padlen = min(max_pad, len/2 - 1), forward pass with initial state zi * x_ext[0],
// backward pass with zi * y[-1], trim.  The biquad cascade is pipelined ACROSS LANES: lane s owns section s and works on
// sample t - s at step t, taking its input from lane s-1 by shuffle -- so the chain per step is one section (3 dependent
// ops), not n_sections.  Each section evaluates scipy's direct-form-II-transposed recurrence with the same operation
// order and without FMA contraction (__dmul_rn/__dadd_rn), so results agree with scipy to rounding.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace bf {

constexpr int kMaxSections = 16;

struct SosCoef {
    double b0[kMaxSections], b1[kMaxSections], b2[kMaxSections], a1[kMaxSections], a2[kMaxSections];
    double zi0[kMaxSections], zi1[kMaxSections];
    int n_sections;
};

// x_ext[i] of the odd extension of run x[s .. s+len-1] with `pad` samples on each side (scipy.signal._arraytools.odd_ext)
__device__ __forceinline__ double odd_ext_at(const double* __restrict__ x, int len, int pad, int i) {
    if (i < pad) return __dsub_rn(__dmul_rn(2.0, x[0]), x[pad - i]);
    if (i < pad + len) return x[i - pad];
    return __dsub_rn(__dmul_rn(2.0, x[len - 1]), x[len - 2 - (i - pad - len)]);
}

// One pass of the cascade over L samples.  `in(i)` is evaluated by all lanes for i = chunk base + lane (coalesced) and
// handed to lane 0 by shuffle; the last section's lane stores out[i] (or calls the sink).
template <typename InF, typename OutF>
__device__ __forceinline__ void sos_pass(const SosCoef& c, int L, double scale, InF in, OutF out) {
    const int lane = threadIdx.x & 31;
    const int S = c.n_sections;
    const bool mine = lane < S;
    const double b0 = mine ? c.b0[lane] : 0.0, b1 = mine ? c.b1[lane] : 0.0, b2 = mine ? c.b2[lane] : 0.0;
    const double a1 = mine ? c.a1[lane] : 0.0, a2 = mine ? c.a2[lane] : 0.0;
    double z0 = mine ? __dmul_rn(c.zi0[lane], scale) : 0.0, z1 = mine ? __dmul_rn(c.zi1[lane], scale) : 0.0;
    double y = 0.0;                                   // my output of the previous step
    double chunk = 0.0;
    for (int t = 0; t < L + S - 1; ++t) {
        if ((t & 31) == 0) chunk = (t + lane < L) ? in(t + lane) : 0.0;
        const double x0 = __shfl_sync(0xffffffffu, chunk, t & 31);
        const double xprev = __shfl_up_sync(0xffffffffu, y, 1);
        const int i = t - lane;                       // sample index this lane works on
        if (mine && i >= 0 && i < L) {
            const double xc = lane == 0 ? x0 : xprev;
            const double yn = __dadd_rn(__dmul_rn(b0, xc), z0);
            z0 = __dadd_rn(__dsub_rn(__dmul_rn(b1, xc), __dmul_rn(a1, yn)), z1);
            z1 = __dsub_rn(__dmul_rn(b2, xc), __dmul_rn(a2, yn));
            y = yn;
            if (lane == S - 1) out(i, yn);
        }
    }
}

__global__ void __launch_bounds__(32) k_bandpass_nanrobust(const double* __restrict__ x, int n, const SosCoef c, int min_len,
                                                           int max_pad, double* __restrict__ y, double* __restrict__ scratch,
                                                           int scratch_stride) {
    const int s = blockIdx.x, lane = threadIdx.x;
    const double* xs = x + (size_t)s * n;
    double* ys = y + (size_t)s * n;
    double* fw = scratch + (size_t)s * scratch_stride;
    const double nanv = nan("");
    for (int i = lane; i < n; i += 32) ys[i] = nanv;
    __syncwarp();
    int run_start = -1;
    for (int base = 0; base < n + 32; base += 32) {          // one extra chunk flushes a run that touches the end
        const int i = base + lane;
        const bool fin = (i < n) && isfinite(xs[i]);
        const unsigned m = __ballot_sync(0xffffffffu, fin);
        for (int b = 0; b < 32; ++b) {                        // warp-uniform walk over the 32 flags
            const bool f = (m >> b) & 1u;
            const int idx = base + b;
            if (f && run_start < 0) run_start = idx;
            if (!f && run_start >= 0) {
                const int len = idx - run_start;
                if (len >= min_len) {
                    const int pad = min(max_pad, len / 2 - 1);
                    const double* xr = xs + run_start;
                    double* yr = ys + run_start;
                    if (pad <= 0) {
                        for (int k = lane; k < len; k += 32) yr[k] = xr[k];
                    } else {
                        const int L = len + 2 * pad;
                        const double x0 = odd_ext_at(xr, len, pad, 0);
                        sos_pass(c, L, x0, [&](int k) { return odd_ext_at(xr, len, pad, k); },
                                 [&](int k, double v) { fw[k] = v; });
                        __syncwarp();
                        __threadfence_block();
                        const double y0 = fw[L - 1];
                        sos_pass(c, L, y0, [&](int k) { return fw[L - 1 - k]; },
                                 [&](int k, double v) {
                                     const int j = L - 1 - k - pad;   // position in the run after the final reversal + trim
                                     if (j >= 0 && j < len) yr[j] = v;
                                 });
                        __syncwarp();
                    }
                }
                run_start = -1;
            }
        }
    }
}

}  // namespace bf
