// Sliding-window 2x2 PCA -> dynamic PC1 (replaces dynamic_pc1_sliding, /root/reference/optical_PCA.py:136-235;
// parallel formulation validated in SURVEY.md Appendix B).  float64 throughout, like the reference.
//
// Three phases, batched over (window configuration, series):
//   A  one thread per window: finite mask, mean-centre, covariance, principal axis, align to `ref`
//      (optical_PCA.py:181-202)
//   B  one CTA per (cfg, series): compact the valid windows, turn the sequential prev_w sign chain
//      (optical_PCA.py:203-205) into a block scan over valid windows.  With a_i the ref-aligned axis of window i and s_i
//      its final sign, the reference tests dot(a_i, s_{i-1} a_{i-1}) < 0: a negative a_i . a_{i-1} flips the running sign,
//      a positive one keeps it, and an exact ZERO leaves a_i as it is (s_i = +1 whatever came before: a reset).  The
//      three cases compose associatively as (reset, parity) pairs.
//   C  one thread per sample: nearest window centre (ties -> later centre, optical_PCA.py:218-225),
//      non-centred projection (optical_PCA.py:227-233)
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace bf {

struct Pc1Cfg {
    int win_n, step_n, K;   // K = number of candidate windows for this configuration
    int win_off;            // offset of this cfg's windows inside the per-series window scratch
};
// Window configurations travel as a kernel parameter (no host-to-device copy from pageable memory in a stream-ordered call).
constexpr int kPc1MaxCfg = 32;
struct Pc1CfgPack { Pc1Cfg c[kPc1MaxCfg]; };

// scratch per (cfg, series): wx[Ktot], wy[Ktot] (aligned axis or compacted+signed axis), cen[Ktot], nvalid
__global__ void k_pc1_windows(const double* __restrict__ vx, const double* __restrict__ vy, int n_series, int n,
                              const Pc1CfgPack cfgs, int n_cfg, int Ktot, double ref_x, double ref_y,
                              int min_samples, double* __restrict__ wx, double* __restrict__ wy,
                              int* __restrict__ valid) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y, c = blockIdx.z;
    const Pc1Cfg cf = cfgs.c[c];
    if (j >= cf.K) return;
    const double* x = vx + (size_t)s * n;
    const double* y = vy + (size_t)s * n;
    const int start = j * cf.step_n;
    int cnt = 0;
    double sx = 0, sy = 0;
    for (int i = 0; i < cf.win_n; ++i) {
        const double a = x[start + i], b = y[start + i];
        if (isfinite(a) && isfinite(b)) { sx += a; sy += b; ++cnt; }
    }
    const size_t o = (size_t)s * Ktot + cf.win_off + j;
    if (cnt < min_samples) { valid[o] = 0; wx[o] = 0; wy[o] = 0; return; }
    const double mx = sx / cnt, my = sy / cnt;
    double sxx = 0, sxy = 0, syy = 0;
    for (int i = 0; i < cf.win_n; ++i) {
        const double a = x[start + i], b = y[start + i];
        if (isfinite(a) && isfinite(b)) {
            const double da = a - mx, db = b - my;
            sxx += da * da; sxy += da * db; syy += db * db;
        }
    }
    // principal axis of [[sxx, sxy], [sxy, syy]]: theta = atan2(2 sxy, sxx - syy) / 2
    double ex, ey;
    if (sxy == 0.0) {
        if (sxx >= syy) { ex = 1.0; ey = 0.0; } else { ex = 0.0; ey = 1.0; }
    } else {
        const double th = 0.5 * atan2(2.0 * sxy, sxx - syy);
        sincos(th, &ey, &ex);
    }
    if (ex * ref_x + ey * ref_y < 0.0) { ex = -ex; ey = -ey; }
    valid[o] = 1; wx[o] = ex; wy[o] = ey;
}

// Block scan helpers (blockDim.x == 1024 or less, power of two)
__device__ __forceinline__ int block_excl_scan_add(int v, int* sh, int* total) {
    const int t = threadIdx.x, nt = blockDim.x;
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < nt; o <<= 1) {
        const int add = (t >= o) ? sh[t - o] : 0;
        __syncthreads();
        sh[t] += add;
        __syncthreads();
    }
    const int incl = sh[t];
    *total = sh[nt - 1];
    __syncthreads();
    return incl - v;
}

// Phase B.  In place: on exit the first nvalid entries of wx/wy/cen hold the signed axes and centres of
// the valid windows in order; nvalid_out[(s, c)] = count.
__global__ void k_pc1_chain(const Pc1CfgPack cfgs, int n_cfg, int Ktot, double* __restrict__ wx,
                            double* __restrict__ wy, const int* __restrict__ valid, double* __restrict__ cwx,
                            double* __restrict__ cwy, int* __restrict__ cen, int* __restrict__ nvalid_out) {
    extern __shared__ int sh[];
    const int s = blockIdx.x, c = blockIdx.y;
    const Pc1Cfg cf = cfgs.c[c];
    const size_t base = (size_t)s * Ktot + cf.win_off;
    const int t = threadIdx.x, nt = blockDim.x;
    const int per = (cf.K + nt - 1) / nt;
    const int lo = min(t * per, cf.K), hi = min(lo + per, cf.K);
    // 1) compaction positions
    int cnt = 0;
    for (int j = lo; j < hi; ++j) cnt += valid[base + j];
    int total;
    int pos = block_excl_scan_add(cnt, sh, &total);
    for (int j = lo; j < hi; ++j) {
        if (valid[base + j]) {
            cwx[base + pos] = wx[base + j];
            cwy[base + pos] = wy[base + j];
            cen[base + pos] = (2 * j * cf.step_n + cf.win_n - 1) / 2;  // (start + end - 1) // 2
            ++pos;
        }
    }
    __syncthreads();
    if (t == 0) nvalid_out[s * n_cfg + c] = total;
    // 2) sign chain over the compacted axes.  Step code of window i from d = a_i . a_{i-1}: 0 keep (d > 0), 1 flip (d < 0),
    //    2 reset (d == 0, or i == 0).  A run of steps composes to (reset << 1 | parity).
    const int per2 = (total + nt - 1) / nt;
    const int lo2 = min(t * per2, total), hi2 = min(lo2 + per2, total);
    auto step_code = [&](int i) -> int {
        if (i == 0) return 2;
        const double d = cwx[base + i] * cwx[base + i - 1] + cwy[base + i] * cwy[base + i - 1];
        return d < 0.0 ? 1 : (d == 0.0 ? 2 : 0);
    };
    auto compose = [](int first, int then) -> int { return (then & 2) ? then : ((first & 2) | ((first ^ then) & 1)); };
    int mine = 0;
    for (int i = lo2; i < hi2; ++i) mine = compose(mine, step_code(i));
    // inclusive Hillis-Steele scan with the (non-commutative) composition; every dot product above is read before anyone
    // flips an axis (the scan's barriers separate the two)
    sh[t] = mine;
    __syncthreads();
    for (int o = 1; o < nt; o <<= 1) {
        const int prev = (t >= o) ? sh[t - o] : 0;
        __syncthreads();
        sh[t] = compose(prev, sh[t]);
        __syncthreads();
    }
    int state = (t > 0) ? sh[t - 1] : 0;                 // composite of everything before this thread's run
    // replay the run: the dots must be taken against the UNFLIPPED neighbours, so read ahead before writing
    double px = 0, py = 0;
    if (lo2 > 0 && lo2 < hi2) { px = cwx[base + lo2 - 1]; py = cwy[base + lo2 - 1]; }
    __syncthreads();
    for (int i = lo2; i < hi2; ++i) {
        const double ax = cwx[base + i], ay = cwy[base + i];
        int code = 2;
        if (i > 0) {
            const double d = ax * px + ay * py;
            code = d < 0.0 ? 1 : (d == 0.0 ? 2 : 0);
        }
        state = compose(state, code);
        px = ax; py = ay;
        if (state & 1) { cwx[base + i] = -ax; cwy[base + i] = -ay; }
    }
}

// Phase C.
__global__ void k_pc1_project(const double* __restrict__ vx, const double* __restrict__ vy, int n_series, int n,
                              const Pc1CfgPack cfgs, int n_cfg, int Ktot, const double* __restrict__ cwx,
                              const double* __restrict__ cwy, const int* __restrict__ cen,
                              const int* __restrict__ nvalid, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y, c = blockIdx.z;
    if (i >= n) return;
    const Pc1Cfg cf = cfgs.c[c];
    const size_t base = (size_t)s * Ktot + cf.win_off;
    const int K = nvalid[s * n_cfg + c];
    double r = nan("");
    if (K > 0) {
        // searchsorted(centres, i, 'left'): first index with centre >= i
        int lo = 0, hi = K;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cen[base + mid] < i) lo = mid + 1; else hi = mid;
        }
        const int j = min(lo, K - 1), j2 = max(j - 1, 0);
        const int pick = (abs(i - cen[base + j2]) < abs(i - cen[base + j])) ? j2 : j;
        const double a = vx[(size_t)s * n + i], b = vy[(size_t)s * n + i];
        if (isfinite(a) && isfinite(b)) r = a * cwx[base + pick] + b * cwy[base + pick];
    }
    out[((size_t)c * n_series + s) * n + i] = r;
}

}  // namespace bf
