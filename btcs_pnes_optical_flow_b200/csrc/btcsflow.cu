// libbtcsflow.so -- C-ABI host side (plan, scheduling, streaming) over the sm_100a kernels.
// Interface contract and the reference file:line each entry point replaces: include/btcsflow.h.
#include "../../include/btcsflow.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "farneback_kernels.cuh"
#include "farneback_fast.cuh"
#include "pc1_kernels.cuh"
#include "bandpass_kernels.cuh"

namespace {

// Grow-only device scratch for the stream-ordered entry points that are not tied to a plan (PC1, band-pass), one block per
// (device, stream) and calling thread: calls on different streams never share a block, calls on one stream are ordered by
// the stream itself.  (cudaMallocAsync/cudaFreeAsync would hand the memory back to the OS at every synchronisation --
// default pool release threshold 0 -- and cost milliseconds per call on a KB-sized problem.)
struct StreamScratch {
    struct Block { int dev; cudaStream_t st; void* buf; size_t cap; };
    std::vector<Block> blocks;
    ~StreamScratch() { /* freed with the context at process exit; a thread's blocks cannot be freed safely while queued work may use them */ }
    cudaError_t get(int dev, cudaStream_t st, size_t need, void** out) {
        for (auto& b : blocks) {
            if (b.dev != dev || b.st != st) continue;
            if (b.cap < need) {
                cudaError_t e = cudaStreamSynchronize(st);            // earlier work on this stream may still use the block
                if (e != cudaSuccess) return e;
                cudaFree(b.buf);
                b.buf = nullptr; b.cap = 0;
                e = cudaMalloc(&b.buf, need + need / 2);
                if (e != cudaSuccess) return e;
                b.cap = need + need / 2;
            }
            *out = b.buf;
            return cudaSuccess;
        }
        if (blocks.size() >= 16) {                                    // bounded: recycle the oldest block
            cudaError_t e = cudaStreamSynchronize(blocks.front().st);
            if (e != cudaSuccess) cudaGetLastError();                 // the stream may have been destroyed: nothing left to wait for
            cudaFree(blocks.front().buf);
            blocks.erase(blocks.begin());
        }
        Block b{dev, st, nullptr, need + need / 2};
        cudaError_t e = cudaMalloc(&b.buf, b.cap);
        if (e != cudaSuccess) return e;
        blocks.push_back(b);
        *out = b.buf;
        return cudaSuccess;
    }
};

// device that owns a device pointer (the stream-ordered entry points run on the data's device, whatever is current)
int device_of(const void* dptr, int* dev) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, dptr) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
        cudaGetLastError();
        return cudaGetDevice(dev) == cudaSuccess ? 0 : 1;
    }
    *dev = at.device;
    return 0;
}

thread_local std::string g_err;
thread_local long long g_launches = 0;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail((int)e_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK()                                                                             \
    do {                                                                                           \
        ++g_launches;                                                                              \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess)                                                                     \
            return fail((int)e_, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int cv_round(double v) { return (int)std::nearbyint(v); }  // round half to even (cvRound)
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// cv2.resize INTER_LINEAR coefficient tables, float32 weights
void resize_table(int dst, int src, std::vector<int>& idx, std::vector<float>& wgt) {
    idx.resize(dst);
    wgt.resize(dst);
    if (dst == src) {
        for (int d = 0; d < dst; ++d) { idx[d] = d; wgt[d] = 0.f; }
        return;
    }
    const double inv = (double)dst / (double)src;
    const double scale = 1.0 / inv;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int i = (int)std::floor(f);
        f -= (float)i;
        if (i < 0) { i = 0; f = 0.f; }
        if (i >= src - 1) { i = src - 1; f = 0.f; }
        idx[d] = i;
        wgt[d] = f;
    }
}

// cv2.getGaussianKernel (float32 taps)
std::vector<float> gaussian_kernel(int ksize, double sigma) {
    std::vector<float> k(ksize);
    if (sigma <= 0 && ksize == 3) { k = {0.25f, 0.5f, 0.25f}; return k; }
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    std::vector<double> d(ksize);
    double s = 0;
    const double c = (ksize - 1) * 0.5;
    for (int i = 0; i < ksize; ++i) { const double x = i - c; d[i] = std::exp(-(x * x) / (2 * sigma * sigma)); s += d[i]; }
    for (int i = 0; i < ksize; ++i) k[i] = (float)(d[i] / s);
    return k;
}

bool invert6(double A[6][6], double inv[6][6]) {
    double a[6][12];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { a[i][j] = A[i][j]; a[i][6 + j] = (i == j); }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (std::fabs(a[piv][c]) < 1e-300) return false;
        if (piv != c) for (int j = 0; j < 12; ++j) std::swap(a[piv][j], a[c][j]);
        const double d = 1.0 / a[c][c];
        for (int j = 0; j < 12; ++j) a[c][j] *= d;
        for (int r = 0; r < 6; ++r) {
            if (r == c) continue;
            const double f = a[r][c];
            if (f != 0) for (int j = 0; j < 12; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) inv[i][j] = a[i][6 + j];
    return true;
}

// FarnebackPrepareGaussian
bool make_poly_coef(int n, double sigma, bf::PolyCoef& pc) {
    if (sigma < 1.1920928955078125e-07) sigma = n * 0.3;
    std::vector<float> g(2 * n + 1), xg(2 * n + 1), xxg(2 * n + 1);
    double s = 0;
    for (int x = -n; x <= n; ++x) { g[x + n] = (float)std::exp(-x * x / (2 * sigma * sigma)); s += g[x + n]; }
    s = 1.0 / s;
    for (int x = -n; x <= n; ++x) {
        g[x + n] = (float)(g[x + n] * s);
        xg[x + n] = (float)(x * g[x + n]);
        xxg[x + n] = (float)(x * x * g[x + n]);
    }
    double G[6][6] = {{0}};
    for (int y = -n; y <= n; ++y)
        for (int x = -n; x <= n; ++x) {
            const float gg = g[y + n] * g[x + n];
            G[0][0] += gg;
            G[1][1] += gg * x * x;
            G[3][3] += gg * x * x * x * x;
            G[5][5] += gg * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    double inv[6][6];
    if (!invert6(G, inv)) return false;
    pc.n = n;
    for (int k = 0; k <= n; ++k) { pc.g[k] = g[n + k]; pc.xg[k] = xg[n + k]; pc.xxg[k] = xxg[n + k]; }
    pc.ig11 = (float)inv[1][1];
    pc.ig03 = (float)inv[0][3];
    pc.ig33 = (float)inv[3][3];
    pc.ig55 = (float)inv[5][5];
    return true;
}

// cv2.resize INTER_AREA coefficient table for one axis (computeResizeAreaTab), CSR form
void area_table(int ssize, int dsize, std::vector<int>& ofs, std::vector<int>& idx, std::vector<float>& wgt) {
    const double scale = (double)ssize / dsize;
    ofs.assign(1, 0); idx.clear(); wgt.clear();
    for (int d = 0; d < dsize; ++d) {
        if (ssize == dsize) { idx.push_back(d); wgt.push_back(1.f); ofs.push_back((int)idx.size()); continue; }
        const double f1 = d * scale, f2 = f1 + scale, cell = std::min(scale, ssize - f1);
        int s1 = (int)std::ceil(f1), s2 = (int)std::floor(f2);
        s2 = std::min(s2, ssize - 1);
        s1 = std::min(s1, s2);
        if (s1 - f1 > 1e-3) { idx.push_back(s1 - 1); wgt.push_back((float)((s1 - f1) / cell)); }
        for (int sx = s1; sx < s2; ++sx) { idx.push_back(sx); wgt.push_back((float)(1.0 / cell)); }
        if (f2 - s2 > 1e-3) { idx.push_back(s2); wgt.push_back((float)(std::min(std::min(f2 - s2, 1.0), cell) / cell)); }
        ofs.push_back((int)idx.size());
    }
}

void make_win_coef(int winsize, int flags, bf::WinCoef& wc) {
    const int m = winsize / 2;
    wc.m = m;
    wc.gauss = (flags & BF_OPTFLOW_FARNEBACK_GAUSSIAN) ? 1 : 0;
    wc.scale = wc.gauss ? 1.f : (float)(1.0 / ((double)winsize * winsize));
    const double sigma = m * 0.3;
    std::vector<float> ker(m + 1);
    ker[0] = 1.f;
    double s = 1.0;
    for (int i = 1; i <= m; ++i) {
        const float t = (float)std::exp(-i * i / (2 * sigma * sigma));
        ker[i] = t;
        s += t * 2;
    }
    s = 1.0 / s;
    for (int i = 0; i <= m; ++i) wc.ker[i] = (float)(ker[i] * s);
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct ScaleInfo {
    int k, w, h, ksize, pitch;
    double scale, sigma;
    size_t plane;          // h * pitch
    // device tables
    int *ix = nullptr, *iy = nullptr;       // image resize: full-res -> this scale
    float *ax = nullptr, *ay = nullptr;
    float* kern = nullptr;                    // pyramid Gaussian taps
    int *fix = nullptr, *fiy = nullptr;     // flow upsample: next-coarser scale -> this scale
    float *fax = nullptr, *fay = nullptr;
    float* I = nullptr;                       // [F][h][pitch]
    float* tmpk = nullptr;                    // horizontal-pass scratch of this level [F][H][pitch] (levels coarser than 0)
    void* R = nullptr;                        // fp32 planes [F][5][h][pitch], or packed fp16 pixels [F][h][pitch] x 16 B
    float2* flow = nullptr;                   // [B][h][pitch]
    // tensor maps for the tile kernel's L2 prefetch of the R blocks (compact plans, box window)
    bf::TileMaps maps[1];
    bool have_maps = false;
};

template <typename T>
cudaError_t upload(const std::vector<T>& v, T** out) {
    cudaError_t e = cudaMalloc((void**)out, std::max<size_t>(v.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

struct bf_plan {
    bf_params prm;
    int W, H, B, F, max_rois, device;
    std::vector<ScaleInfo> sc;  // coarse -> fine
    bf::PolyCoef pc;
    bf::WinCoef wc;
    float* tmp = nullptr;       // [F][H][pitch0]
    void* M[2] = {nullptr, nullptr};   // matrices ping-pong: f32 planes, or fp16 planes on compact plans
    float* axes = nullptr;      // [B][4]
    float* partial = nullptr;   // [B][max_rois][ncta][kRoiVals]
    uint8_t* roi_class = nullptr;   // [max_rois][ncta]: class of every tile of the last launch (k_roi_tile_class), per series call
    int ncta_max = 0;
    size_t bytes = 0;
    // host-staging path
    uint8_t* stage[2] = {nullptr, nullptr};
    float* stage_flow = nullptr;
    uint8_t* pair_in[2] = {nullptr, nullptr};   // bf_flow_pair_host staging (H x W x 4 bytes each)
    float* pair_flow = nullptr;
    // per-call resources of the host series path, two slots used alternately (a call may be queued while the previous one
    // still runs): device axes / masks / result, and a pinned host block the caller's (pageable) axes and masks are copied
    // into so that their upload never blocks the host behind the stream's earlier work
    struct HostSlot {
        double* d_ex = nullptr; double* d_ey = nullptr; int axes_cap = 0;
        uint8_t* d_masks = nullptr; size_t masks_cap = 0;
        float* d_out = nullptr; size_t out_cap = 0;
        uint8_t* pinned = nullptr; size_t pinned_cap = 0;
        long long last_ticket = -1;                 // the call that used this slot last
    } slot[2];
    cudaEvent_t ev_aux = nullptr;                   // axes + masks of the current call have reached the device
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    bool ev_free_rec[2] = {false, false};       // staging buffer b has a pending "expansion done" event (also across calls)
    static constexpr int kTickets = 8;
    cudaEvent_t ev_done[kTickets] = {};         // completion events of bf_flow_series_host_async calls (ring)
    long long next_ticket = 0;
    // OPTFLOW_USE_INITIAL_FLOW: INTER_AREA tables full resolution -> coarsest scale
    int *ia_xofs = nullptr, *ia_xidx = nullptr, *ia_yofs = nullptr, *ia_yidx = nullptr;
    float *ia_xwgt = nullptr, *ia_ywgt = nullptr;
    bool use_fast = true;
    bool r_half = false;        // polynomial coefficients packed in 16 B per pixel, b f32 + A f16 (fast path, uint8 input)
    int sm_count = 148;
    // stage timing (bf_plan_profile): CUDA event pairs on the launch stream, tagged BF_PROF_*
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;   // start/stop pairs
    std::vector<int> prof_tag, prof_np; // per pair of events
    size_t prof_used = 0;               // events in use
    int prof_n[BF_PROF_NTAGS] = {};     // results of the last bf_plan_profile_read
    double prof_ms[BF_PROF_NTAGS] = {};
    long long prof_pairs[BF_PROF_NTAGS] = {};
};

namespace {

template <typename T>
int plan_alloc(bf_plan* p, T** ptr, size_t count) {
    CU(cudaMalloc((void**)ptr, std::max<size_t>(count, 1) * sizeof(T)));
    p->bytes += count * sizeof(T);
    return 0;
}

int validate_params(const bf_params* q, int W, int H) {
    if (!q) return fail(BF_E_INVALID, "params is NULL");
    if (W <= 0 || H <= 0) return fail(BF_E_INVALID, "bad image size %dx%d", W, H);
    if (!(q->pyr_scale < 1.0) || !(q->pyr_scale > 0.0))
        return fail(BF_E_INVALID,
                    "(-215:Assertion failed) prev0.size() == next0.size() && prev0.channels() == next0.channels() "
                    "&& prev0.channels() == 1 && pyrScale_ < 1 (pyr_scale=%g)", q->pyr_scale);
    if (q->levels < 0 || q->iterations < 0) return fail(BF_E_INVALID, "levels/iterations must be >= 0");
    if (q->winsize < 1) return fail(BF_E_INVALID, "winsize must be >= 1");
    if (q->poly_n < 1) return fail(BF_E_INVALID, "poly_n must be >= 1");
    if (q->poly_n > BF_MAX_POLY_N) return fail(BF_E_UNSUPPORTED, "poly_n=%d > %d not supported", q->poly_n, BF_MAX_POLY_N);
    if (q->winsize / 2 > BF_MAX_WIN_HALF) return fail(BF_E_UNSUPPORTED, "winsize=%d too large (max %d)", q->winsize, 2 * BF_MAX_WIN_HALF + 1);
    if (q->flags & ~(BF_OPTFLOW_FARNEBACK_GAUSSIAN | BF_OPTFLOW_USE_INITIAL_FLOW))
        return fail(BF_E_INVALID, "unknown flags 0x%x", q->flags);
    return 0;
}

// ---- launches -----------------------------------------------------------------------------------------

template <typename T>
int launch_pyramid(bf_plan* p, const ScaleInfo& s, const T* src, size_t pitch_bytes, size_t frame_bytes, int nf,
                   float* out, int out_pitch, size_t out_frame_stride, cudaStream_t st) {
    const int tmp_pitch = s.pitch;
    const size_t tmp_stride = (size_t)p->H * tmp_pitch;
    dim3 b(64, 4);
    dim3 g1(cdiv(s.w, 64), cdiv(p->H, 4), nf);
    bf::k_pyr_h<T><<<g1, b, 0, st>>>(src, pitch_bytes, frame_bytes, p->W, p->H, s.w, s.ix, s.ax, s.kern, s.ksize,
                                      p->tmp, tmp_pitch, tmp_stride);
    LAUNCH_CHECK();
    dim3 g2(cdiv(s.w, 64), cdiv(s.h, 4), nf);
    bf::k_pyr_v<<<g2, b, 0, st>>>(p->tmp, tmp_pitch, tmp_stride, p->H, s.w, s.h, s.iy, s.ay, s.kern, s.ksize, out,
                                  out_pitch, out_frame_stride);
    LAUNCH_CHECK();
    return 0;
}

int launch_polyexp(const float* I, int pitch, size_t frame_stride, int w, int h, void* R, size_t plane_stride,
                   size_t slot_stride, int slot0, int nslots, int nf, const bf::PolyCoef& pc, bool allow_fast, bool r_half,
                   cudaStream_t st) {
    if (allow_fast && bf::polyexp_fast_supported(pc.n, pitch) && bf::polyexp_fast_aligned(R, plane_stride, slot_stride, I, frame_stride)) {
        bf::launch_polyexp_fast(I, pitch, frame_stride, w, h, R, plane_stride, slot_stride, slot0, nslots, nf, pc, r_half, st);
        LAUNCH_CHECK();
        return 0;
    }
    if (r_half) return fail(BF_E_UNSUPPORTED, "internal: packed R needs the compile-time polyexp kernel");
    dim3 g(cdiv(w, bf::kPeTW), cdiv(h, bf::kPeTH), nf);
    bf::k_polyexp_generic<<<g, 256, 0, st>>>(I, pitch, frame_stride, w, h, static_cast<float*>(R), plane_stride, slot_stride,
                                             slot0, nslots, pc);
    LAUNCH_CHECK();
    return 0;
}

// Pairs per group in the CTA order of the batched kernels (farneback_common.cuh, decode_cta).  8: frame p+1's coefficients
// (R1 of pair p, R0 of pair p+1) are read from HBM once per group instead of twice -- measured DRAM reads of the finest
// iteration launch 6.29 -> 4.44 GB, of the first update 4.13 -> 2.65 GB (ncu, profiles/r2d_pair_group_dram.txt); the kernels are
// latency-bound, so the time moves by 1-3 % only.  BTCSFLOW_PAIR_GROUP overrides (1 = tile-major order).
int pair_group() {
    const char* e = getenv("BTCSFLOW_PAIR_GROUP");
    const int v = e ? atoi(e) : 8;
    return v >= 1 && v <= 64 ? v : 8;
}

int launch_update(const bf::UpdateArgs& a0, int np, bool r_half, cudaStream_t st) {
    bf::UpdateArgs a = a0;
    a.np = np; a.pair_group = pair_group();
    dim3 b(64, 4);
    const unsigned g = (unsigned)(cdiv(a.w, 64) * cdiv(a.h, 4 * bf::kUpdRows)) * (unsigned)np;
    if (r_half) bf::k_update<true><<<g, b, 0, st>>>(a);
    else bf::k_update<false><<<g, b, 0, st>>>(a);
    LAUNCH_CHECK();
    return 0;
}

// Which kernel runs one blur+solve(+update) iteration: a compile-time-window tile kernel (box or Gaussian, half windows
// 2..16) or the runtime-parameter kernel (BTCSFLOW_NO_FAST=1, other windows, unaligned stage-API buffers).
enum BlurKernel { BK_GENERIC = 0, BK_TILE = 1, BK_GAUSS = 3 };

BlurKernel choose_blur_kernel(const bf::BlurSolveArgs& a, const bf::WinCoef& wc, bool allow_fast) {
    if (!allow_fast || !bf::tile_fast_aligned(a)) return BK_GENERIC;
    if (bf::gauss_fast_supported(wc)) return BK_GAUSS;
    if (bf::box_fast_supported(wc, a.pitch)) return BK_TILE;
    return BK_GENERIC;
}

int blur_solve_ncta(const bf::BlurSolveArgs& a, const bf::WinCoef& wc, bool allow_fast, bool r_half) {
    switch (choose_blur_kernel(a, wc, allow_fast)) {
        case BK_TILE: return bf::box_fast_ncta(a.w, a.h, r_half);
        case BK_GAUSS: return bf::gauss_fast_ncta(a.w, a.h, r_half);
        default: return cdiv(a.w, bf::kBsTW) * cdiv(a.h, bf::kBsTH);
    }
}

int launch_blur_solve(const bf::BlurSolveArgs& a0, const bf::WinCoef& wc, int np, bool allow_fast, bool r_half,
                      cudaStream_t st, const bf::TileMaps* maps = nullptr) {
    bf::BlurSolveArgs a = a0;
    a.np = np; a.pair_group = pair_group();
    switch (choose_blur_kernel(a, wc, allow_fast)) {
        case BK_TILE:
            bf::launch_box_fast(a, wc, np, r_half, st, maps);
            break;
        case BK_GAUSS:
            bf::launch_gauss_fast(a, wc, np, r_half, st);
            break;
        default: {
            dim3 g(cdiv(a.w, bf::kBsTW), cdiv(a.h, bf::kBsTH), np);
            if (r_half) bf::k_blur_solve_generic<true><<<g, 256, 0, st>>>(a, wc);
            else bf::k_blur_solve_generic<false><<<g, 256, 0, st>>>(a, wc);
        }
    }
    LAUNCH_CHECK();
    return 0;
}

// Pyramid + polynomial expansion of nf frames whose first frame index is tf (ring slot = frame % F).
template <typename T>
int expand_frames(bf_plan* p, const T* frames, size_t pitch_bytes, size_t frame_bytes, int tf, int nf,
                  cudaStream_t st) {
    if (p->use_fast && bf::pyr_h_smem_bytes(p->W) <= 160 * 1024) {
        // level 0: exact 3x3 stencil; coarser levels: one multi-level horizontal pass + per-level vertical pass
        bf::PyrHArgs pa{};
        for (auto& s : p->sc) {
            if (s.k == 0) {
                dim3 g(cdiv(p->W, bf::kL0TW), cdiv(p->H, bf::kL0TH), nf);
                bf::k_level0_blur<T><<<g, 128, 0, st>>>(frames, pitch_bytes, frame_bytes, p->W, p->H, s.I, s.pitch, s.plane);
                LAUNCH_CHECK();
            } else if (pa.nlev < bf::kPyrMaxLevels) {
                bf::PyrLevelDesc& d = pa.lv[pa.nlev++];
                d.ix = s.ix; d.ax = s.ax; d.kern = s.kern; d.ksize = s.ksize; d.w = s.w;
                d.tmp = s.tmpk; d.tmp_pitch = s.pitch; d.tmp_frame_stride = (size_t)p->H * s.pitch;
            }
        }
        if (pa.nlev > 0) {
            const size_t smem = bf::pyr_h_smem_bytes(p->W);
            cudaFuncSetAttribute(bf::k_pyr_h_multi<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            bf::k_pyr_h_multi<T><<<dim3(cdiv(p->H, bf::kPyrRows), nf), 256, smem, st>>>(frames, pitch_bytes, frame_bytes, p->W,
                                                                                      p->H, pa);
            LAUNCH_CHECK();
            for (auto& s : p->sc) {
                if (s.k == 0) continue;
                if (bf::pyr_v4_ok(s.tmpk, s.pitch, (size_t)p->H * s.pitch, s.I, s.pitch, s.plane)) {
                    dim3 b(32, 8), g2(cdiv(s.w, 128), cdiv(s.h, 8), nf);
                    bf::k_pyr_v4<<<g2, b, 0, st>>>(s.tmpk, s.pitch, (size_t)p->H * s.pitch, p->H, s.w, s.h, s.iy, s.ay, s.kern,
                                                   s.ksize, s.I, s.pitch, s.plane);
                } else {
                    dim3 b(64, 4), g2(cdiv(s.w, 64), cdiv(s.h, 4), nf);
                    bf::k_pyr_v<<<g2, b, 0, st>>>(s.tmpk, s.pitch, (size_t)p->H * s.pitch, p->H, s.w, s.h, s.iy, s.ay, s.kern,
                                                  s.ksize, s.I, s.pitch, s.plane);
                }
                LAUNCH_CHECK();
            }
        }
    } else {
        for (auto& s : p->sc) {
            int rc = launch_pyramid<T>(p, s, frames, pitch_bytes, frame_bytes, nf, s.I, s.pitch, s.plane, st);
            if (rc) return rc;
        }
    }
    for (auto& s : p->sc) {
        int rc = launch_polyexp(s.I, s.pitch, s.plane, s.w, s.h, s.R, s.plane, p->r_half ? s.plane : 5 * s.plane, tf % p->F,
                                p->F, nf, p->pc, p->use_fast, p->r_half, st);
        if (rc) return rc;
    }
    return 0;
}

int prof_begin(bf_plan* p, int tag, int np, cudaStream_t st) {
    if (!p->prof_on) return 0;
    if (p->prof_used + 2 > p->prof_ev.size()) {
        for (int e = 0; e < 2; ++e) {
            cudaEvent_t ev;
            CU(cudaEventCreate(&ev));
            p->prof_ev.push_back(ev);
        }
        p->prof_tag.push_back(0);
        p->prof_np.push_back(0);
    }
    p->prof_tag[p->prof_used / 2] = tag;
    p->prof_np[p->prof_used / 2] = np;
    CU(cudaEventRecord(p->prof_ev[p->prof_used], st));
    return 0;
}
int prof_end(bf_plan* p, cudaStream_t st) {
    if (!p->prof_on) return 0;
    CU(cudaEventRecord(p->prof_ev[p->prof_used + 1], st));
    p->prof_used += 2;
    return 0;
}

struct RoiCtx {
    bool classes_ready = false;      // p->roi_class holds the tile classes of these masks (computed at the first batch)
    const uint8_t* masks = nullptr;  // [n_roi][H][W]
    int n_roi = 0;
    const double* ex = nullptr;      // device [T][2]
    const double* ey = nullptr;
    float* out = nullptr;            // device [n_roi][T][3]
    int T = 0;
};

// Coarse-to-fine schedule for np pairs whose first frame is t0 (pair q = frames t0+q, t0+q+1).
// flow_out: dense [np][H][W][2] or NULL.  roi: optional ROI reduction into roi->out rows t0+1+q.
// init_flow: OPTFLOW_USE_INITIAL_FLOW -- dense [np][H][W][2] flow the coarsest scale starts from (cv2: INTER_AREA resize
// times the scale); may alias flow_out (read at the coarsest scale, written by the last launch).
int run_pairs(bf_plan* p, int t0, int np, float* flow_out, RoiCtx* roi, cudaStream_t st, const float* init_flow = nullptr) {
    const int I = p->prm.iterations;
    const int nsc = (int)p->sc.size();
    const int slot0 = t0 % p->F;
    const bool want_roi = roi && roi->n_roi > 0;
    if (want_roi) {
        bf::k_axes_to_f32<<<cdiv(np, 64), 64, 0, st>>>(roi->ex, roi->ey, t0 + 1, np, p->axes);
        LAUNCH_CHECK();
    }
    if (I == 0 && init_flow) return fail(BF_E_UNSUPPORTED, "OPTFLOW_USE_INITIAL_FLOW with iterations = 0 is not supported");
    if (I == 0) {
        // cv2 leaves the (zero-initialised, upsampled) flow untouched: the result is identically zero
        const ScaleInfo& s = p->sc.back();
        if (flow_out) CU(cudaMemsetAsync(flow_out, 0, (size_t)np * p->H * p->W * sizeof(float2), st));
        if (want_roi) {
            CU(cudaMemsetAsync(s.flow, 0, (size_t)np * s.plane * sizeof(float2), st));
            bf::BlurSolveArgs a{};
            a.w = s.w; a.h = s.h; a.pitch = s.pitch;
            a.flow = s.flow; a.flow_pitch = s.pitch; a.flow_stride = s.plane;
            a.masks = roi->masks; a.n_roi = roi->n_roi; a.mask_stride = (size_t)p->H * p->W; a.mask_pitch = p->W;
            a.axes = p->axes; a.partial = p->partial;
            dim3 g(cdiv(s.w, bf::kBsTW), cdiv(s.h, bf::kBsTH), np);
            bf::k_roi_from_flow<<<g, 256, 0, st>>>(a);
            LAUNCH_CHECK();
            const int ncta = g.x * g.y;
            bf::k_roi_finalize<<<cdiv(np * roi->n_roi, 4), 128, 0, st>>>(p->partial, np, roi->n_roi, ncta, roi->ex,
                                                                         roi->ey, t0 + 1, roi->out, roi->T);
            LAUNCH_CHECK();
        }
        return 0;
    }
    for (int i = 0; i < nsc; ++i) {
        const ScaleInfo& s = p->sc[i];
        const bool finest = (i == nsc - 1);
        const size_t m_stride = bf::m_pair_bytes(p->r_half, s.w, s.h, s.plane);
        bf::UpdateArgs u{};
        u.R = s.R; u.plane_stride = s.plane; u.slot_stride = p->r_half ? s.plane : 5 * s.plane; u.slot0 = slot0; u.nslots = p->F;
        u.pitch = s.pitch; u.w = s.w; u.h = s.h;
        if (i == 0 && init_flow) {
            dim3 g(cdiv(s.w, 256), s.h, np);
            bf::k_resize_area_flow<<<g, 256, 0, st>>>((const float2*)init_flow, p->W, (size_t)p->H * p->W,
                                                     bf::AreaTab{p->ia_xofs, p->ia_xidx, p->ia_xwgt}, bf::AreaTab{p->ia_yofs, p->ia_yidx, p->ia_ywgt},
                                                     s.w, s.h, (float)s.scale, s.flow, s.pitch, s.plane);
            LAUNCH_CHECK();
            u.flow_mode = 1;
            u.flow = s.flow; u.flow_pitch = s.pitch; u.flow_stride = s.plane;
        } else if (i == 0) {
            u.flow_mode = 0;
        } else {
            const ScaleInfo& c = p->sc[i - 1];
            u.flow_mode = 2;
            u.flow = c.flow; u.flow_pitch = c.pitch; u.flow_stride = c.plane;
            u.ws = c.w; u.hs = c.h; u.mult = (float)(1.0 / p->prm.pyr_scale);
            u.tab = bf::ResizeTab{s.fix, s.fax, s.fiy, s.fay};
        }
        u.M = p->M[0]; u.m_stride = m_stride;
        int rc = prof_begin(p, finest ? BF_PROF_UPDATE : BF_PROF_COARSE, np, st);
        if (rc) return rc;
        rc = launch_update(u, np, p->r_half, st);
        if (rc) return rc;
        if ((rc = prof_end(p, st))) return rc;
        for (int it = 0; it < I; ++it) {
            const bool last = (it == I - 1);
            bf::BlurSolveArgs a{};
            a.M = p->M[it & 1]; a.m_stride = m_stride; a.plane_stride = s.plane; a.pitch = s.pitch; a.w = s.w; a.h = s.h;
            a.R = s.R; a.slot_stride = p->r_half ? s.plane : 5 * s.plane; a.slot0 = slot0; a.nslots = p->F;
            if (!last) {
                a.Mout = p->M[(it + 1) & 1];
            } else if (!finest) {
                a.flow = s.flow; a.flow_pitch = s.pitch; a.flow_stride = s.plane;
            } else {
                if (flow_out) {
                    a.flow = (float2*)flow_out; a.flow_pitch = p->W; a.flow_stride = (size_t)p->H * p->W;
                }
                if (want_roi) {
                    a.masks = roi->masks; a.n_roi = roi->n_roi; a.mask_stride = (size_t)p->H * p->W; a.mask_pitch = p->W;
                    a.axes = p->axes; a.partial = p->partial;
                    // tile classes of the masks (tile kernels only), once per series call
                    const BlurKernel bk = choose_blur_kernel(a, p->wc, p->use_fast);
                    if (bk == BK_TILE || bk == BK_GAUSS) {
                        const int th = bk == BK_TILE ? bf::box_tile_th(p->r_half) : bf::gauss_tile_th(p->r_half);
                        const int nbx = cdiv(s.w, bf::kFbTW), nby = cdiv(s.h, th);
                        if (!roi->classes_ready) {
                            bf::k_roi_tile_class<<<dim3(cdiv(nbx * nby, 4), roi->n_roi), 128, 0, st>>>(roi->masks, roi->n_roi, p->W, p->H, bf::kFbTW, th,
                                                                                                      nbx, nby, p->roi_class);
                            LAUNCH_CHECK();
                            roi->classes_ready = true;
                        }
                        a.roi_class = p->roi_class;
                    }
                }
            }
            if ((rc = prof_begin(p, finest ? (last ? BF_PROF_ITER_LAST : BF_PROF_ITER_UPDATE) : BF_PROF_COARSE, np, st))) return rc;
            rc = launch_blur_solve(a, p->wc, np, p->use_fast, p->r_half, st, s.have_maps ? &s.maps[0] : nullptr);
            if (rc) return rc;
            if ((rc = prof_end(p, st))) return rc;
            if (last && finest && want_roi) {
                const int ncta = blur_solve_ncta(a, p->wc, p->use_fast, p->r_half);
                bf::k_roi_finalize<<<cdiv(np * roi->n_roi, 4), 128, 0, st>>>(p->partial, np, roi->n_roi, ncta, roi->ex,
                                                                             roi->ey, t0 + 1, roi->out, roi->T);
                LAUNCH_CHECK();
            }
        }
    }
    return 0;
}

int check_plan(const bf_plan* p) {
    if (!p) return fail(BF_E_INVALID, "plan is NULL");
    return 0;
}

}  // namespace

// =========================================================================================================
// C-ABI
// =========================================================================================================
extern "C" {

const char* bf_last_error(void) { return g_err.c_str(); }
const char* bf_version(void) { return "btcsflow 0.1 (sm_100a)"; }
long long bf_launch_count(void) { return g_launches; }
void bf_launch_count_reset(void) { g_launches = 0; }

int bf_device_sm(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        cudaGetLastError();
        return fail(BF_E_NODEVICE, "no usable CUDA device %d (this library has no CPU fallback)", device);
    }
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) return fail(BF_E_NODEVICE, "cudaGetDeviceProperties failed");
    return pr.major * 10 + pr.minor;
}

int bf_plan_create(const bf_params* params, int width, int height, int max_pairs, int max_rois, int device,
                   bf_plan** out) {
    return bf_plan_create_ex(params, width, height, max_pairs, max_rois, device, 0, out);
}

int bf_plan_create_ex(const bf_params* params, int width, int height, int max_pairs, int max_rois, int device,
                      unsigned flags, bf_plan** out) {
    if (!out) return fail(BF_E_INVALID, "out is NULL");
    *out = nullptr;
    int rc = validate_params(params, width, height);
    if (rc) return rc;
    if (max_pairs < 1 || max_rois < 0) return fail(BF_E_INVALID, "max_pairs must be >= 1 and max_rois >= 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(BF_E_NODEVICE, "no usable CUDA device %d (this library has no CPU fallback)", device);
    }
    DeviceGuard dg(device);
    if (!dg.ok) return fail(BF_E_NODEVICE, "cudaSetDevice(%d) failed", device);

    bf_plan* p = new bf_plan();
    p->prm = *params;
    p->W = width; p->H = height; p->B = max_pairs; p->F = max_pairs + 1; p->max_rois = max_rois; p->device = device;
    if (!make_poly_coef(params->poly_n, params->poly_sigma, p->pc)) {
        delete p;
        return fail(BF_E_INVALID, "singular moment matrix for poly_n=%d poly_sigma=%g", params->poly_n, params->poly_sigma);
    }
    make_win_coef(params->winsize, params->flags, p->wc);
    {
        const char* nofast = getenv("BTCSFLOW_NO_FAST");
        p->use_fast = !(nofast && nofast[0] == '1');
        const char* rs = getenv("BTCSFLOW_R_STORAGE");
        const bool want_f32 = (flags & BF_PLAN_EXACT_F32) || (rs && strcmp(rs, "f32") == 0);
        // packed fp16 coefficients need the compile-time polyexp kernels (poly_n 5 / 7) and bounded (uint8) input
        // (and frames of at least 4 x 2 pixels: the packed gather reads a 2 x 2 footprint at immediate offsets)
        // Gaussian windows below 8 pixels (sigma <= 0.9) concentrate the weight on one or two samples: in flat regions of the
        // attenuated border ring the fp16 G terms become subnormal and the compact result moves by up to 0.4 px there
        // (tools/window_sweep.py).  Those windows are cheap anyway: they take the all-fp32 storage.
        const bool tiny_gauss = (params->flags & BF_OPTFLOW_FARNEBACK_GAUSSIAN) && params->winsize < 8;
        p->r_half = p->use_fast && !want_f32 && !tiny_gauss && (params->poly_n == 5 || params->poly_n == 7) && width >= 4 && height >= 2;
        cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    }

    // scale selection
    int k = 0;
    double scv = 1.0;
    for (; k < params->levels; ++k) {
        scv *= params->pyr_scale;
        if (width * scv < 32 || height * scv < 32) break;
    }
    const int L = k;
    if (L + 1 > BF_MAX_SCALES) { delete p; return fail(BF_E_UNSUPPORTED, "too many scales (%d)", L + 1); }
    for (int kk = L; kk >= 0; --kk) {
        ScaleInfo s{};
        double scale = 1.0;
        for (int i = 0; i < kk; ++i) scale *= params->pyr_scale;
        s.k = kk; s.scale = scale;
        s.sigma = (1.0 / scale - 1.0) * 0.5;
        s.ksize = std::max(cv_round(s.sigma * 5) | 1, 3);
        s.w = cv_round(width * scale);
        s.h = cv_round(height * scale);
        s.pitch = round_up(s.w, 32);
        s.plane = (size_t)s.h * s.pitch;
        p->sc.push_back(s);
    }
    auto cleanup_fail = [&](int code) { bf_plan_destroy(p); return code; };
    for (size_t i = 0; i < p->sc.size(); ++i) {
        ScaleInfo& s = p->sc[i];
        std::vector<int> idx; std::vector<float> wgt;
        resize_table(s.w, width, idx, wgt);
        if (upload(idx, &s.ix) != cudaSuccess || upload(wgt, &s.ax) != cudaSuccess) return cleanup_fail(fail(2, "table upload failed"));
        resize_table(s.h, height, idx, wgt);
        if (upload(idx, &s.iy) != cudaSuccess || upload(wgt, &s.ay) != cudaSuccess) return cleanup_fail(fail(2, "table upload failed"));
        std::vector<float> kern = gaussian_kernel(s.ksize, s.sigma);
        if (upload(kern, &s.kern) != cudaSuccess) return cleanup_fail(fail(2, "table upload failed"));
        if (i > 0) {
            const ScaleInfo& c = p->sc[i - 1];
            resize_table(s.w, c.w, idx, wgt);
            if (upload(idx, &s.fix) != cudaSuccess || upload(wgt, &s.fax) != cudaSuccess) return cleanup_fail(fail(2, "table upload failed"));
            resize_table(s.h, c.h, idx, wgt);
            if (upload(idx, &s.fiy) != cudaSuccess || upload(wgt, &s.fay) != cudaSuccess) return cleanup_fail(fail(2, "table upload failed"));
        }
        if ((rc = plan_alloc(p, &s.I, (size_t)p->F * s.plane))) return cleanup_fail(rc);
        if (s.k > 0 && (rc = plan_alloc(p, &s.tmpk, (size_t)p->F * height * s.pitch))) return cleanup_fail(rc);
        {
            uint8_t* rbuf = nullptr;
            const size_t rbytes = (size_t)p->F * s.plane * (p->r_half ? sizeof(uint4) : 5 * sizeof(float));
            if ((rc = plan_alloc(p, &rbuf, rbytes))) return cleanup_fail(rc);
            cudaMemset(rbuf, 0, rbytes);
            s.R = rbuf;
        }
        if ((rc = plan_alloc(p, &s.flow, (size_t)p->B * s.plane))) return cleanup_fail(rc);
        // rows [h, pitch) padding must hold finite values for vectorised kernels: zero everything once
        cudaMemset(s.I, 0, (size_t)p->F * s.plane * sizeof(float));
        cudaMemset(s.flow, 0, (size_t)p->B * s.plane * sizeof(float2));
    }
    if (params->flags & BF_OPTFLOW_USE_INITIAL_FLOW) {
        const ScaleInfo& c = p->sc.front();
        std::vector<int> ofs, idx; std::vector<float> wgt;
        area_table(width, c.w, ofs, idx, wgt);
        if (upload(ofs, &p->ia_xofs) != cudaSuccess || upload(idx, &p->ia_xidx) != cudaSuccess || upload(wgt, &p->ia_xwgt) != cudaSuccess)
            return cleanup_fail(fail(2, "table upload failed"));
        area_table(height, c.h, ofs, idx, wgt);
        if (upload(ofs, &p->ia_yofs) != cudaSuccess || upload(idx, &p->ia_yidx) != cudaSuccess || upload(wgt, &p->ia_ywgt) != cudaSuccess)
            return cleanup_fail(fail(2, "table upload failed"));
    }
    const ScaleInfo& fine = p->sc.back();
    size_t tmp_elems = 0, m_bytes = 0;
    for (auto& s : p->sc) {
        tmp_elems = std::max(tmp_elems, (size_t)p->F * height * s.pitch);
        m_bytes = std::max(m_bytes, (size_t)p->B * bf::m_pair_bytes(p->r_half, s.w, s.h, s.plane));
    }
    if ((rc = plan_alloc(p, &p->tmp, tmp_elems))) return cleanup_fail(rc);
    for (int i = 0; i < 2; ++i) {
        uint8_t* mbuf = nullptr;
        if ((rc = plan_alloc(p, &mbuf, m_bytes))) return cleanup_fail(rc);
        cudaMemset(mbuf, 0, m_bytes);
        p->M[i] = mbuf;
    }
    // Tensor maps for the tile kernel's L2 prefetch (compact plans; BTCSFLOW_TMAP=0 keeps the per-row requests).
    {
        const char* e = getenv("BTCSFLOW_TMAP");
        if (p->r_half && p->use_fast && bf::box_fast_supported(p->wc, 32) && !(e && e[0] == '0')) {
            const int th = bf::box_tile_th(true);
            for (auto& s : p->sc) {
                if (!bf::tile_fast_shape(s.w, s.h)) continue;
                s.have_maps = bf::encode_tile_maps(&s.maps[0], s.R, s.w, s.h, s.pitch, s.plane, p->F, th);
            }
        }
    }
    if ((rc = plan_alloc(p, &p->axes, (size_t)p->B * 4))) return cleanup_fail(rc);
    p->ncta_max = std::max(cdiv(fine.w, bf::kBsTW) * cdiv(fine.h, bf::kBsTH), cdiv(fine.w, bf::kFbTW) * cdiv(fine.h, 16));
    if ((rc = plan_alloc(p, &p->partial, (size_t)p->B * std::max(max_rois, 1) * p->ncta_max * bf::kRoiVals))) return cleanup_fail(rc);
    if ((rc = plan_alloc(p, &p->roi_class, (size_t)std::max(max_rois, 1) * p->ncta_max))) return cleanup_fail(rc);
    if (cudaDeviceSynchronize() != cudaSuccess) return cleanup_fail(fail(2, "plan initialisation failed: %s", cudaGetErrorString(cudaGetLastError())));
    *out = p;
    return 0;
}

int bf_plan_destroy(bf_plan* p) {
    if (!p) return 0;
    DeviceGuard dg(p->device);
    for (auto& s : p->sc) {
        cudaFree(s.ix); cudaFree(s.iy); cudaFree(s.ax); cudaFree(s.ay); cudaFree(s.kern);
        cudaFree(s.fix); cudaFree(s.fiy); cudaFree(s.fax); cudaFree(s.fay);
        cudaFree(s.I); cudaFree(s.R); cudaFree(s.flow); cudaFree(s.tmpk);
    }
    cudaFree(p->tmp); cudaFree(p->M[0]); cudaFree(p->M[1]); cudaFree(p->axes); cudaFree(p->partial); cudaFree(p->roi_class);
    cudaFree(p->ia_xofs); cudaFree(p->ia_xidx); cudaFree(p->ia_xwgt); cudaFree(p->ia_yofs); cudaFree(p->ia_yidx); cudaFree(p->ia_ywgt);
    cudaFree(p->stage[0]); cudaFree(p->stage[1]); cudaFree(p->stage_flow);
    cudaFree(p->pair_in[0]); cudaFree(p->pair_in[1]); cudaFree(p->pair_flow);
    for (auto& hs : p->slot) {
        cudaFree(hs.d_ex); cudaFree(hs.d_ey); cudaFree(hs.d_masks); cudaFree(hs.d_out);
        if (hs.pinned) cudaFreeHost(hs.pinned);
    }
    if (p->ev_aux) cudaEventDestroy(p->ev_aux);
    for (int i = 0; i < 2; ++i) {
        if (p->ev_copied[i]) cudaEventDestroy(p->ev_copied[i]);
        if (p->ev_free[i]) cudaEventDestroy(p->ev_free[i]);
    }
    for (auto ev : p->ev_done) if (ev) cudaEventDestroy(ev);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    for (auto ev : p->prof_ev) cudaEventDestroy(ev);
    delete p;
    return 0;
}

int bf_plan_profile(bf_plan* p, int enable) {
    int rc = check_plan(p);
    if (rc) return rc;
    p->prof_on = enable != 0;
    return 0;
}

int bf_plan_profile_read(bf_plan* p, int* n_launches, double* total_ms, long long* pair_iterations) {
    int rc = check_plan(p);
    if (rc) return rc;
    DeviceGuard dg(p->device);
    for (int t = 0; t < BF_PROF_NTAGS; ++t) { p->prof_n[t] = 0; p->prof_ms[t] = 0; p->prof_pairs[t] = 0; }
    for (size_t i = 0; i + 1 < p->prof_used; i += 2) {
        CU(cudaEventSynchronize(p->prof_ev[i + 1]));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, p->prof_ev[i], p->prof_ev[i + 1]));
        const int t = p->prof_tag[i / 2];
        p->prof_n[t] += 1;
        p->prof_ms[t] += ms;
        p->prof_pairs[t] += p->prof_np[i / 2];
    }
    if (n_launches) *n_launches = p->prof_n[BF_PROF_ITER_UPDATE] + p->prof_n[BF_PROF_ITER_LAST];
    if (total_ms) *total_ms = p->prof_ms[BF_PROF_ITER_UPDATE] + p->prof_ms[BF_PROF_ITER_LAST];
    if (pair_iterations) *pair_iterations = p->prof_pairs[BF_PROF_ITER_UPDATE] + p->prof_pairs[BF_PROF_ITER_LAST];
    p->prof_used = 0;
    return 0;
}

int bf_plan_profile_tag(const bf_plan* p, int tag, int* n_launches, double* total_ms, long long* pairs) {
    if (!p) return fail(BF_E_INVALID, "plan is NULL");
    if (tag < 0 || tag >= BF_PROF_NTAGS) return fail(BF_E_INVALID, "unknown profile tag %d", tag);
    if (n_launches) *n_launches = p->prof_n[tag];
    if (total_ms) *total_ms = p->prof_ms[tag];
    if (pairs) *pairs = p->prof_pairs[tag];
    return 0;
}

size_t bf_plan_workspace_bytes(const bf_plan* p) { return p ? p->bytes : 0; }
int bf_plan_num_scales(const bf_plan* p) { return p ? (int)p->sc.size() : BF_E_INVALID; }
int bf_plan_coeff_storage(const bf_plan* p) { return p ? (p->r_half ? 16 : 32) : BF_E_INVALID; }

int bf_plan_scale_info(const bf_plan* p, int i, int* w, int* h, int* ksize, double* sigma, int* pitch) {
    int rc = check_plan(p);
    if (rc) return rc;
    if (i < 0 || i >= (int)p->sc.size()) return fail(BF_E_INVALID, "scale index %d out of range", i);
    const ScaleInfo& s = p->sc[i];
    if (w) *w = s.w;
    if (h) *h = s.h;
    if (ksize) *ksize = s.ksize;
    if (sigma) *sigma = s.sigma;
    if (pitch) *pitch = s.pitch;
    return 0;
}

int bf_flow_pair(bf_plan* p, const void* prev, const void* next, int dtype, size_t pitch_bytes, float* flow_out,
                 void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    if (!prev || !next || !flow_out) return fail(BF_E_INVALID, "NULL image/flow pointer");
    if (dtype != BF_DTYPE_U8 && dtype != BF_DTYPE_F32) return fail(BF_E_INVALID, "dtype must be BF_DTYPE_U8 or BF_DTYPE_F32");
    const size_t esz = dtype == BF_DTYPE_U8 ? 1 : 4;
    if (pitch_bytes < (size_t)p->W * esz) return fail(BF_E_INVALID, "pitch_bytes smaller than a row");
    if (dtype == BF_DTYPE_F32 && p->r_half)
        return fail(BF_E_UNSUPPORTED, "float32 input needs a plan created with BF_PLAN_EXACT_F32 (fp16 coefficient storage "
                                      "is only range-safe for uint8 frames)");
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const void* fr[2] = {prev, next};
    for (int i = 0; i < 2; ++i) {
        if (dtype == BF_DTYPE_U8) rc = expand_frames<uint8_t>(p, (const uint8_t*)fr[i], pitch_bytes, 0, i, 1, st);
        else rc = expand_frames<float>(p, (const float*)fr[i], pitch_bytes, 0, i, 1, st);
        if (rc) return rc;
    }
    // OPTFLOW_USE_INITIAL_FLOW: flow_out is in/out, as the `flow` argument of the cv2 call
    return run_pairs(p, 0, 1, flow_out, nullptr, st, (p->prm.flags & BF_OPTFLOW_USE_INITIAL_FLOW) ? flow_out : nullptr);
}

int bf_flow_pair_host(bf_plan* p, const void* prev, const void* next, int dtype, size_t pitch_bytes, float* flow_out,
                      void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    if (!prev || !next || !flow_out) return fail(BF_E_INVALID, "NULL image/flow pointer");
    if (dtype != BF_DTYPE_U8 && dtype != BF_DTYPE_F32) return fail(BF_E_INVALID, "dtype must be BF_DTYPE_U8 or BF_DTYPE_F32");
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = dtype == BF_DTYPE_U8 ? 1 : 4;
    if (pitch_bytes < (size_t)p->W * esz) return fail(BF_E_INVALID, "pitch_bytes smaller than a row");
    const size_t row_bytes = (size_t)p->W * esz;
    const size_t flow_bytes = (size_t)p->H * p->W * sizeof(float2);
    if (!p->pair_flow) {                         // plan-owned staging: no allocation on the per-pair hot path
        for (int i = 0; i < 2; ++i)
            if ((rc = plan_alloc(p, &p->pair_in[i], (size_t)p->H * p->W * 4))) return rc;
        if ((rc = plan_alloc(p, &p->pair_flow, (size_t)p->H * p->W * 2))) return rc;
    }
    CU(cudaMemcpy2DAsync(p->pair_in[0], row_bytes, prev, pitch_bytes, row_bytes, p->H, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpy2DAsync(p->pair_in[1], row_bytes, next, pitch_bytes, row_bytes, p->H, cudaMemcpyHostToDevice, st));
    if (p->prm.flags & BF_OPTFLOW_USE_INITIAL_FLOW) CU(cudaMemcpyAsync(p->pair_flow, flow_out, flow_bytes, cudaMemcpyHostToDevice, st));
    rc = bf_flow_pair(p, p->pair_in[0], p->pair_in[1], dtype, row_bytes, p->pair_flow, stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(flow_out, p->pair_flow, flow_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int bf_flow_series(bf_plan* p, const uint8_t* frames, int T, const double* ex, const double* ey,
                   const uint8_t* roi_masks, int n_roi, float* out, float* flow_out, void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    if (p->prm.flags & BF_OPTFLOW_USE_INITIAL_FLOW) return fail(BF_E_UNSUPPORTED, "OPTFLOW_USE_INITIAL_FLOW applies to the pair call (bf_flow_pair*): a series has no caller-provided flow per pair");
    if (T < 1 || !frames) return fail(BF_E_INVALID, "need frames and T >= 1");
    if (n_roi < 0 || n_roi > p->max_rois) return fail(BF_E_INVALID, "n_roi=%d exceeds plan max_rois=%d", n_roi, p->max_rois);
    if (n_roi > 0 && (!roi_masks || !out || !ex || !ey)) return fail(BF_E_INVALID, "ROI reduction needs masks, axes and out");
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t fb = (size_t)p->H * p->W;
    if (n_roi > 0) {
        const int n = n_roi * T * 3;
        bf::k_fill_nan<<<cdiv(n, 256), 256, 0, st>>>(out, n);
        LAUNCH_CHECK();
    }
    RoiCtx roi;
    roi.masks = roi_masks; roi.n_roi = n_roi; roi.ex = ex; roi.ey = ey; roi.out = out; roi.T = T;
    for (int t0 = 0; t0 < T - 1; t0 += p->B) {
        const int np = std::min(p->B, T - 1 - t0);
        const int tf = (t0 == 0) ? 0 : t0 + 1;  // frame t0 of a later batch is already expanded (ring)
        const int nf = t0 + np - tf + 1;
        if ((rc = prof_begin(p, BF_PROF_EXPAND, nf, st))) return rc;
        rc = expand_frames<uint8_t>(p, frames + (size_t)tf * fb, (size_t)p->W, fb, tf, nf, st);
        if (rc) return rc;
        if ((rc = prof_end(p, st))) return rc;
        rc = run_pairs(p, t0, np, flow_out ? flow_out + (size_t)t0 * fb * 2 : nullptr, n_roi > 0 ? &roi : nullptr, st);
        if (rc) return rc;
    }
    return 0;
}

int bf_flow_series_host_async(bf_plan* p, const uint8_t* frames, int T, const double* ex, const double* ey,
                              const uint8_t* roi_masks, int n_roi, float* out, float* flow_out, void* stream,
                              long long* ticket) {
    int rc = check_plan(p);
    if (rc) return rc;
    if (!ticket) return fail(BF_E_INVALID, "ticket is NULL");
    if (p->prm.flags & BF_OPTFLOW_USE_INITIAL_FLOW) return fail(BF_E_UNSUPPORTED, "OPTFLOW_USE_INITIAL_FLOW applies to the pair call (bf_flow_pair*): a series has no caller-provided flow per pair");
    if (T < 1 || !frames) return fail(BF_E_INVALID, "need frames and T >= 1");
    if (n_roi < 0 || n_roi > p->max_rois) return fail(BF_E_INVALID, "n_roi=%d exceeds plan max_rois=%d", n_roi, p->max_rois);
    if (n_roi > 0 && (!roi_masks || !out || !ex || !ey)) return fail(BF_E_INVALID, "ROI reduction needs masks, axes and out");
    DeviceGuard dg(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t fb = (size_t)p->H * p->W;
    // lazily created staging resources
    if (!p->copy_stream) {
        CU(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&p->ev_copied[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_free[i], cudaEventDisableTiming));
            if ((rc = plan_alloc(p, &p->stage[i], (size_t)p->F * fb))) return rc;
        }
    }
    if (flow_out && !p->stage_flow) {
        if ((rc = plan_alloc(p, &p->stage_flow, (size_t)p->B * fb * 2))) return rc;
    }
    RoiCtx roi;
    // Chunks of up to B pairs, double-buffered: chunk c+1 travels on the copy stream while chunk c computes.  The chunks
    // grow 8 -> 24 -> B: compute starts after ~8 frames have arrived, and since a pair takes ~3.5x longer to compute than a
    // frame takes to cross PCIe, each chunk's copy is hidden behind the previous (3x smaller) chunk's compute.
    int ramp[4] = {8, 24, 0, 0};                       // BTCSFLOW_HOST_CHUNKS="a,b,..." overrides the ramp (experiments)
    if (const char* e = getenv("BTCSFLOW_HOST_CHUNKS")) {
        int k = 0;
        ramp[0] = ramp[1] = 0;
        for (const char* q = e; *q && k < 4; ++k) {
            ramp[k] = atoi(q);
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
    }
    auto chunk_cap = [&](int chunk) { return (chunk < 4 && ramp[chunk] > 0) ? std::min(p->B, ramp[chunk]) : p->B; };
    auto issue_copy = [&](int chunk, int t0) -> int {          // frames of chunk `chunk` (pairs t0 .. t0+np-1) -> staging
        const int np = std::min(chunk_cap(chunk), T - 1 - t0);
        const int tf = (t0 == 0) ? 0 : t0 + 1;
        const int nf = t0 + np - tf + 1;
        const int b = chunk & 1;
        // the staging buffer must have been consumed by the expansion of the chunk that used it last (this call's chunk - 2,
        // or the tail of an earlier asynchronous call)
        if (p->ev_free_rec[b]) CU(cudaStreamWaitEvent(p->copy_stream, p->ev_free[b], 0));
        CU(cudaMemcpyAsync(p->stage[b], frames + (size_t)tf * fb, (size_t)nf * fb, cudaMemcpyHostToDevice, p->copy_stream));
        CU(cudaEventRecord(p->ev_copied[b], p->copy_stream));
        return 0;
    };
    // Calls on one plan share its workspace: whatever stream this call is given, it starts after the previous call's work.
    if (p->next_ticket > 0) CU(cudaStreamWaitEvent(st, p->ev_done[(p->next_ticket - 1) % bf_plan::kTickets], 0));
    // The first chunk's frames are queued on the copy stream BEFORE the ROI masks and axes: the first frame should not wait.
    if (T > 1 && (rc = issue_copy(0, 0))) return rc;
    bool aux_pending = false;
    if (n_roi > 0) {
        bf_plan::HostSlot& hs = p->slot[p->next_ticket & 1];
        // the slot's buffers may still be read by the call that used it last (two calls back): wait for that one only
        if (hs.last_ticket >= 0) CU(cudaEventSynchronize(p->ev_done[hs.last_ticket % bf_plan::kTickets]));
        if (hs.axes_cap < T) {
            cudaFree(hs.d_ex); cudaFree(hs.d_ey); hs.d_ex = hs.d_ey = nullptr;
            CU(cudaMalloc((void**)&hs.d_ex, (size_t)T * 2 * sizeof(double)));
            CU(cudaMalloc((void**)&hs.d_ey, (size_t)T * 2 * sizeof(double)));
            hs.axes_cap = T;
        }
        const size_t mb = (size_t)n_roi * fb;
        if (hs.masks_cap < mb) {
            cudaFree(hs.d_masks); hs.d_masks = nullptr;
            CU(cudaMalloc(&hs.d_masks, mb));
            hs.masks_cap = mb;
        }
        const size_t ob = (size_t)n_roi * T * 3;
        if (hs.out_cap < ob) {
            cudaFree(hs.d_out); hs.d_out = nullptr;
            CU(cudaMalloc((void**)&hs.d_out, ob * sizeof(float)));
            hs.out_cap = ob;
        }
        const size_t ab = (size_t)T * 2 * sizeof(double);
        if (hs.pinned_cap < 2 * ab + mb) {
            if (hs.pinned) cudaFreeHost(hs.pinned);
            hs.pinned = nullptr;
            CU(cudaMallocHost((void**)&hs.pinned, 2 * ab + mb));
            hs.pinned_cap = 2 * ab + mb;
        }
        memcpy(hs.pinned, ex, ab);
        memcpy(hs.pinned + ab, ey, ab);
        memcpy(hs.pinned + 2 * ab, roi_masks, mb);
        CU(cudaMemcpyAsync(hs.d_ex, hs.pinned, ab, cudaMemcpyHostToDevice, p->copy_stream));
        CU(cudaMemcpyAsync(hs.d_ey, hs.pinned + ab, ab, cudaMemcpyHostToDevice, p->copy_stream));
        CU(cudaMemcpyAsync(hs.d_masks, hs.pinned + 2 * ab, mb, cudaMemcpyHostToDevice, p->copy_stream));
        if (!p->ev_aux) CU(cudaEventCreateWithFlags(&p->ev_aux, cudaEventDisableTiming));
        CU(cudaEventRecord(p->ev_aux, p->copy_stream));
        aux_pending = true;
        const int n = (int)ob;
        bf::k_fill_nan<<<cdiv(n, 256), 256, 0, st>>>(hs.d_out, n);
        LAUNCH_CHECK();
        hs.last_ticket = p->next_ticket;
        roi.masks = hs.d_masks; roi.n_roi = n_roi; roi.ex = hs.d_ex; roi.ey = hs.d_ey; roi.out = hs.d_out; roi.T = T;
    }
    int chunk = 0;
    for (int t0 = 0; t0 < T - 1; ++chunk) {
        const int np = std::min(chunk_cap(chunk), T - 1 - t0);
        const int tf = (t0 == 0) ? 0 : t0 + 1;
        const int nf = t0 + np - tf + 1;
        const int b = chunk & 1;
        // the NEXT chunk's copy is queued before this chunk's ~45 launches so that the copy stream stays a chunk ahead
        if (t0 + np < T - 1 && (rc = issue_copy(chunk + 1, t0 + np))) return rc;
        CU(cudaStreamWaitEvent(st, p->ev_copied[b], 0));
        if ((rc = prof_begin(p, BF_PROF_EXPAND, nf, st))) return rc;
        rc = expand_frames<uint8_t>(p, p->stage[b], (size_t)p->W, fb, tf, nf, st);
        if (rc) return rc;
        if ((rc = prof_end(p, st))) return rc;
        CU(cudaEventRecord(p->ev_free[b], st));
        p->ev_free_rec[b] = true;
        if (aux_pending) { CU(cudaStreamWaitEvent(st, p->ev_aux, 0)); aux_pending = false; }
        rc = run_pairs(p, t0, np, flow_out ? p->stage_flow : nullptr, n_roi > 0 ? &roi : nullptr, st);
        if (rc) return rc;
        if (flow_out)
            CU(cudaMemcpyAsync(flow_out + (size_t)t0 * fb * 2, p->stage_flow, (size_t)np * fb * 2 * sizeof(float),
                               cudaMemcpyDeviceToHost, st));
        t0 += np;
    }
    if (n_roi > 0)
        CU(cudaMemcpyAsync(out, roi.out, (size_t)n_roi * T * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    const int slot = (int)(p->next_ticket % bf_plan::kTickets);
    if (!p->ev_done[slot]) CU(cudaEventCreateWithFlags(&p->ev_done[slot], cudaEventDisableTiming));
    CU(cudaEventRecord(p->ev_done[slot], st));
    *ticket = p->next_ticket++;
    return 0;
}

int bf_flow_series_wait(bf_plan* p, long long ticket) {
    int rc = check_plan(p);
    if (rc) return rc;
    if (ticket < 0 || ticket >= p->next_ticket) return fail(BF_E_INVALID, "unknown ticket %lld", ticket);
    // if the slot was reused by a later call, that call was queued behind this one: waiting for it is sufficient
    DeviceGuard dg(p->device);
    CU(cudaEventSynchronize(p->ev_done[ticket % bf_plan::kTickets]));
    return 0;
}

int bf_flow_series_host(bf_plan* p, const uint8_t* frames, int T, const double* ex, const double* ey,
                        const uint8_t* roi_masks, int n_roi, float* out, float* flow_out, void* stream) {
    long long ticket = -1;
    int rc = bf_flow_series_host_async(p, frames, T, ex, ey, roi_masks, n_roi, out, flow_out, stream, &ticket);
    if (rc) return rc;
    DeviceGuard dg(p->device);
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    CU(cudaStreamSynchronize(p->copy_stream));
    return 0;
}

// ---- PC1 ------------------------------------------------------------------------------------------------

int bf_pc1_sliding_batched(const double* vx, const double* vy, int n_series, int n, const int* win_n,
                           const int* step_n, int n_cfg, double ref_x, double ref_y, int min_samples,
                           double* pc1_out, void* stream) {
    if (!vx || !vy || !pc1_out || !win_n || !step_n) return fail(BF_E_INVALID, "NULL pointer");
    if (n_series < 1 || n < 0 || n_cfg < 1) return fail(BF_E_INVALID, "bad sizes");
    if (min_samples < 1) return fail(BF_E_INVALID, "min_samples must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return 0;
    for (int c = 0; c < n_cfg; ++c)
        if (win_n[c] < 1 || step_n[c] < 1) return fail(BF_E_INVALID, "win_n and step_n must be >= 1");
    static thread_local StreamScratch scratch;
    int dev = 0;
    if (device_of(vx, &dev)) return fail(BF_E_NODEVICE, "no current CUDA device");
    DeviceGuard dg(dev);
    if (!dg.ok) return fail(BF_E_NODEVICE, "cudaSetDevice(%d) failed", dev);
    // window configurations travel as kernel parameters, kPc1MaxCfg per round of launches
    for (int c0 = 0; c0 < n_cfg; c0 += bf::kPc1MaxCfg) {
        const int nc = std::min(bf::kPc1MaxCfg, n_cfg - c0);
        bf::Pc1CfgPack pack{};
        int Ktot = 0, Kmax = 0;
        for (int c = 0; c < nc; ++c) {
            // optical_PCA.py:171-172,181: fewer than MIN_SAMPLES samples, or n < win_n -> no windows -> all NaN
            int K = 0;
            if (n >= min_samples && n >= win_n[c0 + c]) K = (n - win_n[c0 + c]) / step_n[c0 + c] + 1;
            pack.c[c] = bf::Pc1Cfg{win_n[c0 + c], step_n[c0 + c], K, Ktot};
            Ktot += K;
            Kmax = std::max(Kmax, K);
        }
        const size_t nw = (size_t)n_series * std::max(Ktot, 1);
        const size_t need = 4 * nw * sizeof(double) + (2 * nw + (size_t)n_series * nc) * sizeof(int);
        void* buf = nullptr;
        CU(scratch.get(dev, st, need, &buf));
        double* d_w = static_cast<double*>(buf);                                           // wx, wy, cwx, cwy
        int* d_i = reinterpret_cast<int*>(d_w + 4 * nw);                                   // valid, cen, nvalid
        double *wx = d_w, *wy = d_w + nw, *cwx = d_w + 2 * nw, *cwy = d_w + 3 * nw;
        int *valid = d_i, *cen = d_i + nw, *nvalid = d_i + 2 * nw;
        if (Kmax > 0) {
            dim3 gA(cdiv(Kmax, 128), n_series, nc);
            bf::k_pc1_windows<<<gA, 128, 0, st>>>(vx, vy, n_series, n, pack, nc, Ktot, ref_x, ref_y, min_samples, wx, wy, valid);
            LAUNCH_CHECK();
        }
        dim3 gB(n_series, nc);
        bf::k_pc1_chain<<<gB, 1024, 1024 * sizeof(int), st>>>(pack, nc, Ktot, wx, wy, valid, cwx, cwy, cen, nvalid);
        LAUNCH_CHECK();
        dim3 gC(cdiv(n, 256), n_series, nc);
        bf::k_pc1_project<<<gC, 256, 0, st>>>(vx, vy, n_series, n, pack, nc, Ktot, cwx, cwy, cen, nvalid,
                                               pc1_out + (size_t)c0 * n_series * n);
        LAUNCH_CHECK();
    }
    return 0;
}

int bf_pc1_sliding(const double* vx, const double* vy, int n, int win_n, int step_n, double ref_x, double ref_y,
                   int min_samples, double* pc1_out, void* stream) {
    return bf_pc1_sliding_batched(vx, vy, 1, n, &win_n, &step_n, 1, ref_x, ref_y, min_samples, pc1_out, stream);
}

int bf_pc1_sliding_host(const double* vx, const double* vy, int n, int win_n, int step_n, double ref_x, double ref_y,
                        int min_samples, double* pc1_out) {
    if (!vx || !vy || !pc1_out) return fail(BF_E_INVALID, "NULL pointer");
    if (n <= 0) return n == 0 ? 0 : fail(BF_E_INVALID, "n < 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(BF_E_NODEVICE, "no usable CUDA device (this library has no CPU fallback)");
    }
    double* d = nullptr;
    CU(cudaMalloc((void**)&d, (size_t)3 * n * sizeof(double)));
    int rc = 0;
    cudaError_t e = cudaMemcpy(d, vx, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + n, vy, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = fail((int)e, "H2D failed: %s", cudaGetErrorString(e));
    if (!rc) rc = bf_pc1_sliding(d, d + n, n, win_n, step_n, ref_x, ref_y, min_samples, d + 2 * n, nullptr);
    if (!rc) {
        e = cudaMemcpy(pc1_out, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail((int)e, "D2H failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

// ---- BGR -> gray -----------------------------------------------------------------------------------------

int bf_bgr2gray(const uint8_t* bgr, int n_frames, int width, int height, size_t in_pitch_bytes, uint8_t* gray,
                size_t out_pitch_bytes, void* stream) {
    if (!bgr || !gray || n_frames < 0 || width <= 0 || height <= 0) return fail(BF_E_INVALID, "bad arguments");
    if (in_pitch_bytes < (size_t)width * 3 || out_pitch_bytes < (size_t)width) return fail(BF_E_INVALID, "pitch smaller than a row");
    if (n_frames == 0) return 0;
    dim3 g(cdiv(cdiv(width, 4), 256), height, n_frames);
    bf::k_bgr2gray<<<g, 256, 0, (cudaStream_t)stream>>>(bgr, in_pitch_bytes, in_pitch_bytes * height, width, height, gray,
                                                        out_pitch_bytes, out_pitch_bytes * height);
    LAUNCH_CHECK();
    return 0;
}

// ---- band-pass -------------------------------------------------------------------------------------------

int bf_sosfilt_zi(const double* sos, int n_sections, double* zi) {
    if (!sos || !zi || n_sections < 1) return fail(BF_E_INVALID, "bad arguments");
    // scipy.signal.sosfilt_zi: per-section lfilter_zi (steady state of the step response), scaled by the DC gain of the
    // sections before it
    double scale = 1.0;
    for (int s = 0; s < n_sections; ++s) {
        const double* q = sos + 6 * s;
        const double a0 = q[3];
        if (a0 == 0.0) return fail(BF_E_INVALID, "sos[%d].a0 is zero", s);
        const double b0 = q[0] / a0, b1 = q[1] / a0, b2 = q[2] / a0, a1 = q[4] / a0, a2 = q[5] / a0;
        const double Bsum = (b1 - a1 * b0) + (b2 - a2 * b0);
        const double z0 = Bsum / ((1.0 + a1) + a2);
        const double z1 = (1.0 + a1) * z0 - (b1 - a1 * b0);
        zi[2 * s] = scale * z0;
        zi[2 * s + 1] = scale * z1;
        scale *= (q[0] + q[1] + q[2]) / (q[3] + q[4] + q[5]);
    }
    return 0;
}

int bf_bandpass_nanrobust(const double* x, int n_series, int n, const double* sos, const double* zi, int n_sections,
                          double* y, void* stream) {
    if (!x || !y || !sos) return fail(BF_E_INVALID, "NULL pointer");
    if (n_series < 1 || n < 0) return fail(BF_E_INVALID, "bad sizes");
    if (n_sections < 1 || n_sections > bf::kMaxSections) return fail(BF_E_UNSUPPORTED, "1..%d second-order sections supported", bf::kMaxSections);
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    bf::SosCoef c{};
    c.n_sections = n_sections;
    std::vector<double> zloc(2 * n_sections);
    if (!zi) {
        int rc = bf_sosfilt_zi(sos, n_sections, zloc.data());
        if (rc) return rc;
        zi = zloc.data();
    }
    for (int s = 0; s < n_sections; ++s) {
        const double* q = sos + 6 * s;
        if (q[3] == 0.0) return fail(BF_E_INVALID, "sos[%d].a0 is zero", s);
        c.b0[s] = q[0] / q[3]; c.b1[s] = q[1] / q[3]; c.b2[s] = q[2] / q[3]; c.a1[s] = q[4] / q[3]; c.a2[s] = q[5] / q[3];
        c.zi0[s] = zi[2 * s]; c.zi1[s] = zi[2 * s + 1];
    }
    // same length rules as the reference: sos_required_padlen = 3 * (2 * n_sections) (optical_PCA.py:74-80, 107, 114)
    const int max_pad = 3 * (2 * n_sections), min_len = max_pad + 1;
    static thread_local StreamScratch scratch;
    int dev = 0;
    if (device_of(x, &dev)) return fail(BF_E_NODEVICE, "no current CUDA device");
    DeviceGuard dg(dev);
    if (!dg.ok) return fail(BF_E_NODEVICE, "cudaSetDevice(%d) failed", dev);
    const int stride = n + 2 * max_pad + 8;
    const size_t need = (size_t)n_series * stride;
    struct { double* buf; } sc{nullptr};
    CU(scratch.get(dev, st, need * sizeof(double), (void**)&sc.buf));
    bf::k_bandpass_nanrobust<<<n_series, 32, 0, st>>>(x, n, c, min_len, max_pad, y, sc.buf, stride);
    LAUNCH_CHECK();
    return 0;
}

// ---- stage-level entry points ---------------------------------------------------------------------------

int bf_stage_level_image(bf_plan* p, const void* frame, int dtype, size_t pitch_bytes, int scale_index, float* out,
                         void* stream) {
    int rc = check_plan(p);
    if (rc) return rc;
    if (!frame || !out) return fail(BF_E_INVALID, "NULL pointer");
    if (scale_index < 0 || scale_index >= (int)p->sc.size()) return fail(BF_E_INVALID, "scale index out of range");
    DeviceGuard dg(p->device);
    const ScaleInfo& s = p->sc[scale_index];
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == BF_DTYPE_U8) return launch_pyramid<uint8_t>(p, s, (const uint8_t*)frame, pitch_bytes, 0, 1, out, s.w, (size_t)s.w * s.h, st);
    if (dtype == BF_DTYPE_F32) return launch_pyramid<float>(p, s, (const float*)frame, pitch_bytes, 0, 1, out, s.w, (size_t)s.w * s.h, st);
    return fail(BF_E_INVALID, "bad dtype");
}

int bf_stage_poly_exp(const float* image, int w, int h, int poly_n, double poly_sigma, float* R_planes, void* stream) {
    if (!image || !R_planes || w <= 0 || h <= 0) return fail(BF_E_INVALID, "bad arguments");
    if (poly_n < 1 || poly_n > BF_MAX_POLY_N) return fail(BF_E_UNSUPPORTED, "poly_n out of range");
    bf::PolyCoef pc;
    if (!make_poly_coef(poly_n, poly_sigma, pc)) return fail(BF_E_INVALID, "singular moment matrix");
    const char* nofast = getenv("BTCSFLOW_NO_FAST");
    const bool fast = !(nofast && nofast[0] == '1');
    return launch_polyexp(image, w, (size_t)w * h, w, h, R_planes, (size_t)w * h, (size_t)5 * w * h, 0, 1, 1, pc, fast, false,
                          (cudaStream_t)stream);
}

int bf_stage_update_matrices(const float* R0, const float* R1, const float* flow, int w, int h, float* M, void* stream) {
    if (!R0 || !R1 || !flow || !M || w <= 0 || h <= 0) return fail(BF_E_INVALID, "bad arguments");
    // R0 and R1 are separate allocations: express R1 as "slot 1" relative to R0 only when contiguous;
    // otherwise run with a two-slot view through pointer difference (must be a multiple of a float).
    const size_t plane = (size_t)w * h;
    bf::UpdateArgs u{};
    const ptrdiff_t diff = R1 - R0;
    u.R = R0; u.plane_stride = plane; u.slot_stride = (size_t)diff; u.slot0 = 0; u.nslots = 2;
    u.pitch = w; u.w = w; u.h = h;
    u.flow_mode = 1; u.flow = (const float2*)flow; u.flow_pitch = w; u.flow_stride = 0;
    u.M = M; u.m_stride = 0;
    return launch_update(u, 1, false, (cudaStream_t)stream);
}

int bf_stage_blur_solve(const float* M, int w, int h, int winsize, int flags, float* flow, void* stream) {
    if (!M || !flow || w <= 0 || h <= 0 || winsize < 1) return fail(BF_E_INVALID, "bad arguments");
    if (winsize / 2 > BF_MAX_WIN_HALF) return fail(BF_E_UNSUPPORTED, "winsize too large");
    bf::WinCoef wc;
    make_win_coef(winsize, flags, wc);
    bf::BlurSolveArgs a{};
    a.M = M; a.m_stride = 0; a.plane_stride = (size_t)w * h; a.pitch = w; a.w = w; a.h = h;
    a.flow = (float2*)flow; a.flow_pitch = w; a.flow_stride = 0;
    const char* nofast = getenv("BTCSFLOW_NO_FAST");
    const bool fast = !(nofast && nofast[0] == '1');
    return launch_blur_solve(a, wc, 1, fast, false, (cudaStream_t)stream);
}

int bf_stage_upsample_flow(const float* flow_in, int ws, int hs, int w, int h, float mult, float* flow_out, void* stream) {
    if (!flow_in || !flow_out || ws <= 0 || hs <= 0 || w <= 0 || h <= 0) return fail(BF_E_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int> ix, iy; std::vector<float> ax, ay;
    resize_table(w, ws, ix, ax);
    resize_table(h, hs, iy, ay);
    int *dix = nullptr, *diy = nullptr; float *dax = nullptr, *day = nullptr;
    CU(cudaMalloc((void**)&dix, w * sizeof(int)));
    CU(cudaMalloc((void**)&diy, h * sizeof(int)));
    CU(cudaMalloc((void**)&dax, w * sizeof(float)));
    CU(cudaMalloc((void**)&day, h * sizeof(float)));
    CU(cudaMemcpyAsync(dix, ix.data(), w * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(diy, iy.data(), h * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dax, ax.data(), w * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(day, ay.data(), h * sizeof(float), cudaMemcpyHostToDevice, st));
    bf::UpdateArgs u{};
    u.w = w; u.h = h; u.flow_mode = 2;
    u.flow = (const float2*)flow_in; u.flow_pitch = ws; u.flow_stride = 0; u.ws = ws; u.hs = hs; u.mult = mult;
    u.tab = bf::ResizeTab{dix, dax, diy, day};
    u.flow_out = (float2*)flow_out; u.flow_out_pitch = w; u.flow_out_stride = 0;
    int rc = launch_update(u, 1, false, st);
    CU(cudaStreamSynchronize(st));  // the host tables above must outlive the async copies
    cudaFree(dix); cudaFree(diy); cudaFree(dax); cudaFree(day);
    return rc;
}

}  // extern "C"

#ifdef BF_TRACE
// Debug builds only (not declared in include/btcsflow.h): point the phase trace of k_blur_solve_box at a device buffer of
// 8 x uint64 per CTA of the largest launch, or at nullptr to stop tracing.
extern "C" int bf_debug_trace_set(void* dev_buf) {
    unsigned long long* p = static_cast<unsigned long long*>(dev_buf);
    return cudaMemcpyToSymbol(bf::bf_trace_buf, &p, sizeof(p)) == cudaSuccess ? 0 : -1;
}
#endif
