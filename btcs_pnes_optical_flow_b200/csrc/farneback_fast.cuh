// Specialised (compile-time poly_n) polynomial expansion, pyramid and colour-conversion kernels.
#pragma once
#include <cstdlib>
#include <cstring>

#include "farneback_kernels.cuh"
#include "farneback_tile.cuh"

namespace bf {

// ---------------------------------------------------------------------------------------------------
// k_polyexp<N>: separable polynomial expansion with compile-time poly_n.  Tile 128 x 32, 256 threads.
//   vertical pass   thread per (column, 8-row group): coalesced 32-bit loads straight from global (rows clamped =
//                   replicate), the 8+2N values live in registers, taps are compile-time constants-bank reads
//   horizontal pass thread per 4 consecutive outputs: conflict-free LDS.128, 128-bit coalesced plane stores
// ---------------------------------------------------------------------------------------------------
template <int N>
struct FastPeCfg {
    // HALO = N exactly: a 4-output group then reads the 2N + 4 values it needs from float4 chunk 0 onwards, all of them
    // LDS.128.  With the halo rounded up to 8 the first and last value sat alone in their chunks and were fetched by scalar
    // LDS with a 16-byte lane stride (4-way bank conflicts, a quarter of the kernel's shared-memory wavefronts; ncu).
    static constexpr int HALO = N;
    static constexpr int NCOL = kFbTW + 2 * HALO;
    static constexpr int VP = (NCOL + 3) / 4 * 4 + 4;
    static constexpr int RG = 4, RPG = kFbTH / RG;             // row groups, rows per group
    static constexpr int NCH = (HALO + N + 3) / 4 + 1;
    static constexpr size_t SMEM = (size_t)3 * kFbTH * VP * sizeof(float);
};

// two adjacent 16-byte pixels as one 256-bit store; p must be 32-byte aligned
__device__ __forceinline__ void store_px2(uint4* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
                 "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// Horizontal pass of the polynomial expansion from the shared vertical-pass result + the 5-coefficient store.
template <int N, bool RH>
__device__ __forceinline__ void polyexp_horizontal_store(const float* smem, int x0, int y0, int f, int pitch, int w, int h,
                                                         void* __restrict__ Rv, size_t plane_stride, size_t slot_stride,
                                                         int slot0, int nslots, const PolyCoef& pc) {
    using C = FastPeCfg<N>;
    const int tid = threadIdx.x;
    const int g = tid & 31, rb = tid >> 5;
    const int slot = (slot0 + f) % nslots;
    float* Rb = RH ? nullptr : static_cast<float*>(Rv) + (size_t)slot * slot_stride;
    uint4* Rh = RH ? static_cast<uint4*>(Rv) + (size_t)slot * slot_stride : nullptr;
    const bool wide = (reinterpret_cast<uintptr_t>(Rh) & 31) == 0;      // pitch and x are multiples of 4 pixels: rows stay 32-byte aligned
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const int r = rb + 8 * k;
        const int y = y0 + r, x = x0 + 4 * g;
        float vv[3][4 * C::NCH];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4* vp = reinterpret_cast<const float4*>(smem + ((size_t)c * kFbTH + r) * C::VP) + g;
#pragma unroll
            for (int i = 0; i < C::NCH; ++i) {
                const float4 t = vp[i];
                vv[c][4 * i] = t.x; vv[c][4 * i + 1] = t.y; vv[c][4 * i + 2] = t.z; vv[c][4 * i + 3] = t.w;
            }
        }
        float o[5][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = C::HALO + j;
            float b1 = vv[0][ci] * pc.g[0], b3 = vv[1][ci] * pc.g[0], b5 = vv[2][ci] * pc.g[0];
            float b2 = 0.f, b4 = 0.f, b6 = 0.f;
#pragma unroll
            for (int t = 1; t <= N; ++t) {
                const float p0 = vv[0][ci + t], m0 = vv[0][ci - t];
                const float p1 = vv[1][ci + t], m1 = vv[1][ci - t];
                const float p2 = vv[2][ci + t], m2 = vv[2][ci - t];
                const float tg = p0 + m0;
                b1 = fmaf(tg, pc.g[t], b1);
                b4 = fmaf(tg, pc.xxg[t], b4);
                b2 = fmaf(p0 - m0, pc.xg[t], b2);
                b3 = fmaf(p1 + m1, pc.g[t], b3);
                b6 = fmaf(p1 - m1, pc.xg[t], b6);
                b5 = fmaf(p2 + m2, pc.g[t], b5);
            }
            o[0][j] = b3 * pc.ig11;
            o[1][j] = b2 * pc.ig11;
            o[2][j] = fmaf(b1, pc.ig03, b5 * pc.ig33);
            o[3][j] = fmaf(b1, pc.ig03, b4 * pc.ig33);
            o[4][j] = b6 * pc.ig55;
        }
        if (RH) {
            if (y < h) {
                uint4* op = Rh + (size_t)y * pitch + x;
                if (x + 3 < w && wide) {
                    // the thread's 4 pixels are 64 contiguous bytes: two 256-bit stores (sm_100a) touch each 128-byte line once
                    // per instruction pair instead of once per pixel (lane stride 64 B)
                    store_px2(op, pack_r(o[0][0], o[1][0], o[2][0], o[3][0], o[4][0]), pack_r(o[0][1], o[1][1], o[2][1], o[3][1], o[4][1]));
                    store_px2(op + 2, pack_r(o[0][2], o[1][2], o[2][2], o[3][2], o[4][2]), pack_r(o[0][3], o[1][3], o[2][3], o[3][3], o[4][3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (x + j < w) op[j] = pack_r(o[0][j], o[1][j], o[2][j], o[3][j], o[4][j]);
                }
            }
        } else if (y < h && x < w) {
            float* op = Rb + (size_t)y * pitch + x;
            if (x + 3 < w) {
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    *reinterpret_cast<float4*>(op + c * plane_stride) = make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
            } else {
#pragma unroll
                for (int c = 0; c < 5; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (x + j < w) op[c * plane_stride + j] = o[c][j];
            }
        }
    }
}

template <int N, bool RH>
__global__ void __launch_bounds__(256, 4) k_polyexp(const float* __restrict__ I, int pitch, size_t frame_stride, int w,
                                                    int h, void* __restrict__ Rv, size_t plane_stride,
                                                    size_t slot_stride, int slot0, int nslots, const PolyCoef pc) {
    using C = FastPeCfg<N>;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFbTW, y0 = blockIdx.y * kFbTH, f = blockIdx.z;
    const float* img = I + (size_t)f * frame_stride;
    for (int task = tid; task < C::RG * C::NCOL; task += 256) {
        const int rg = task / C::NCOL, col = task - rg * C::NCOL;
        const int gx = min(max(x0 - C::HALO + col, 0), w - 1);
        const int yb = y0 + rg * C::RPG - N;
        float v[C::RPG + 2 * N];
#pragma unroll
        for (int i = 0; i < C::RPG + 2 * N; ++i) v[i] = __ldg(img + (size_t)min(max(yb + i, 0), h - 1) * pitch + gx);
        float* dst = smem + (rg * C::RPG) * C::VP + col;
#pragma unroll
        for (int j = 0; j < C::RPG; ++j) {
            const float c = v[j + N];
            float t0 = c * pc.g[0], t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                const float up = v[j + N - k], dn = v[j + N + k];
                const float pp = up + dn;
                t0 = fmaf(pc.g[k], pp, t0);
                t1 = fmaf(pc.xg[k], dn - up, t1);
                t2 = fmaf(pc.xxg[k], pp, t2);
            }
            dst[j * C::VP] = t0;
            dst[(kFbTH + j) * C::VP] = t1;
            dst[(2 * kFbTH + j) * C::VP] = t2;
        }
    }
    __syncthreads();
    polyexp_horizontal_store<N, RH>(smem, x0, y0, f, pitch, w, h, Rv, plane_stride, slot_stride, slot0, nslots, pc);
}

// ---------------------------------------------------------------------------------------------------
// Pyramid.  Every level is made from the FULL-RESOLUTION frame: GaussianBlur(REFLECT_101) then bilinear
// resize.  Scale 1.0 (level k = 0) always uses the fixed [1 2 1]/4 taps and no resize: k_level0_blur does that 3x3
// stencil from a shared uint8->float tile (for uint8 input every product and partial sum is exactly representable, so
// the result is bit-identical to cv2's row-then-column order).  Coarser levels: k_pyr_h_multi evaluates the horizontal
// blur only at the columns the resize samples, for ALL coarser levels from one shared copy of each source row.
// ---------------------------------------------------------------------------------------------------
constexpr int kL0TW = 128, kL0TH = 32, kL0R = 8;          // tile of a 128-thread block; rows walked by one thread

// One thread = 4 adjacent pixels x kL0R rows walking down: a row costs one 32-bit load (+ the two neighbour bytes), the
// horizontally filtered rows slide through registers, every output row is one 128-bit store.  ~12 instructions per pixel;
// the shared-tile version it replaces spent 78 (index arithmetic and reflection tests per loaded element; ncu, profiles/).
template <typename T>
__global__ void __launch_bounds__(128) k_level0_blur(const T* __restrict__ src, size_t src_pitch_bytes, size_t src_frame_bytes,
                                                     int W, int H, float* __restrict__ out, int out_pitch,
                                                     size_t out_frame_stride) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x = (blockIdx.x * 32 + lane) * 4;
    const int yb = blockIdx.y * kL0TH + wid * kL0R;
    const int f = blockIdx.z;
    if (x >= W || yb >= H) return;
    const char* base = (const char*)src + (size_t)f * src_frame_bytes;
    float* obase = out + (size_t)f * out_frame_stride;
    // interior threads of aligned uint8 frames fetch their 4 pixels as one word
    const bool word = sizeof(T) == 1 && x >= 1 && x + 4 < W &&
                      ((src_pitch_bytes | src_frame_bytes | reinterpret_cast<uintptr_t>(src)) & 3) == 0;
    const bool vec_out = x + 3 < W && (out_pitch & 3) == 0 && (out_frame_stride & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    auto hrow = [&](int yy, float hr[4]) {                                    // row filter first, like cv2
        const T* row = (const T*)(base + (size_t)reflect101(min(yy, H), H) * src_pitch_bytes);
        float v[6];
        if (word) {
            const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(row + x));
            v[0] = load_px(row + x - 1);
            v[1] = (float)q.x; v[2] = (float)q.y; v[3] = (float)q.z; v[4] = (float)q.w;
            v[5] = load_px(row + x + 4);
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) v[i] = load_px(row + reflect101(min(x - 1 + i, W), W));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) hr[j] = v[j] * 0.25f + v[j + 1] * 0.5f + v[j + 2] * 0.25f;
    };
    float h0[4], h1[4], h2[4];
    hrow(yb - 1, h0);
    hrow(yb, h1);
#pragma unroll
    for (int k = 0; k < kL0R; ++k) {
        const int y = yb + k;
        if (y >= H) break;
        hrow(y + 1, h2);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = h0[j] * 0.25f + h1[j] * 0.5f + h2[j] * 0.25f;
        float* op = obase + (size_t)y * out_pitch + x;
        if (vec_out) {
            *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j < W) op[j] = o[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { h0[j] = h1[j]; h1[j] = h2[j]; }
    }
}

// k_polyexp_l0<N, RH, T>: pyramid level 0 and its polynomial expansion in one kernel.  The uint8 (or float) frame tile
// goes to shared memory, the exact 3x3 [1 2 1]/4 REFLECT_101 blur (scale 1) is evaluated into a second shared
// tile at REPLICATE-clamped coordinates (what the expansion's borders need), and the two separable passes of
// k_polyexp run from there.  Saves the level-0 image round trip through HBM (4 B/px written + read) and one launch.
template <int N>
struct FusedPeCfg {
    using P = FastPeCfg<N>;
    static constexpr int BROWS = kFbTH + 2 * N;                 // blurred tile rows
    static constexpr int BP = P::NCOL + 4;                      // blurred tile pitch
    static constexpr int UROWS = BROWS + 2, UP = P::NCOL + 2 + 2;   // raw tile (one more ring for the 3x3 blur)
    static constexpr int T_FLOATS = 3 * kFbTH * P::VP;          // vertical-pass result; the raw tile aliases it
    static constexpr size_t SMEM = (size_t)(T_FLOATS + BROWS * BP) * sizeof(float);
    static_assert(UROWS * UP <= T_FLOATS, "raw tile must fit in the aliased region");
};

template <int N, bool RH, typename T>
__global__ void __launch_bounds__(256, 2) k_polyexp_l0(const T* __restrict__ src, size_t src_pitch_bytes, size_t src_frame_bytes,
                                                       int pitch, int w, int h, void* __restrict__ Rv, size_t plane_stride,
                                                       size_t slot_stride, int slot0, int nslots, const PolyCoef pc) {
    using C = FastPeCfg<N>;
    using Fz = FusedPeCfg<N>;
    extern __shared__ __align__(16) float smem[];
    float* tbuf = smem;                                   // [3][TH][VP]   (raw tile U[UROWS][UP] lives here first)
    float* U = smem;
    float* Bt = smem + Fz::T_FLOATS;                      // [BROWS][BP]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFbTW, y0 = blockIdx.y * kFbTH, f = blockIdx.z;
    const char* base = (const char*)src + (size_t)f * src_frame_bytes;
    // raw tile: U[uy][ux] = frame[reflect101(y0 - N - 1 + uy)][reflect101(x0 - HALO - 1 + ux)] (coordinates limited to
    // [-1, size] first: anything further out is never used by an in-image blurred sample)
    for (int e = tid; e < Fz::UROWS * (C::NCOL + 2); e += 256) {
        const int uy = e / (C::NCOL + 2), ux = e - uy * (C::NCOL + 2);
        const int gy = reflect101(min(max(y0 - N - 1 + uy, -1), h), h), gx = reflect101(min(max(x0 - C::HALO - 1 + ux, -1), w), w);
        U[uy * Fz::UP + ux] = load_px((const T*)(base + (size_t)gy * src_pitch_bytes) + gx);
    }
    __syncthreads();
    // blurred tile at replicate-clamped coordinates: tile index ranges that fall inside the image
    const int ty_lo = max(0, N - y0), ty_hi = min(Fz::BROWS - 1, (h - 1) - (y0 - N));
    const int tx_lo = max(0, C::HALO - x0), tx_hi = min(C::NCOL - 1, (w - 1) - (x0 - C::HALO));
    for (int e = tid; e < Fz::BROWS * C::NCOL; e += 256) {
        const int ty = e / C::NCOL, tx = e - ty * C::NCOL;
        const int tyc = min(max(ty, ty_lo), ty_hi), txc = min(max(tx, tx_lo), tx_hi);
        const float* q = U + tyc * Fz::UP + txc;                               // top-left of the 3x3 neighbourhood
        float hr[3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) hr[dy] = q[dy * Fz::UP] * 0.25f + q[dy * Fz::UP + 1] * 0.5f + q[dy * Fz::UP + 2] * 0.25f;
        Bt[ty * Fz::BP + tx] = hr[0] * 0.25f + hr[1] * 0.5f + hr[2] * 0.25f;   // row filter first, like cv2
    }
    __syncthreads();
    // vertical pass from the blurred tile (already border-replicated)
    for (int task = tid; task < C::RG * C::NCOL; task += 256) {
        const int rg = task / C::NCOL, col = task - rg * C::NCOL;
        const float* bsrc = Bt + (rg * C::RPG) * Fz::BP + col;
        float v[C::RPG + 2 * N];
#pragma unroll
        for (int i = 0; i < C::RPG + 2 * N; ++i) v[i] = bsrc[i * Fz::BP];
        float* dst = tbuf + (rg * C::RPG) * C::VP + col;
#pragma unroll
        for (int j = 0; j < C::RPG; ++j) {
            const float c = v[j + N];
            float t0 = c * pc.g[0], t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                const float up = v[j + N - k], dn = v[j + N + k];
                const float pp = up + dn;
                t0 = fmaf(pc.g[k], pp, t0);
                t1 = fmaf(pc.xg[k], dn - up, t1);
                t2 = fmaf(pc.xxg[k], pp, t2);
            }
            dst[j * C::VP] = t0;
            dst[(kFbTH + j) * C::VP] = t1;
            dst[(2 * kFbTH + j) * C::VP] = t2;
        }
    }
    __syncthreads();
    polyexp_horizontal_store<N, RH>(tbuf, x0, y0, f, pitch, w, h, Rv, plane_stride, slot_stride, slot0, nslots, pc);
}

// BGR -> gray exactly like cv2.cvtColor(COLOR_BGR2GRAY) on uint8 (/root/reference/optical_flow.py:227; SURVEY section 8 row
// f-2): 15-bit fixed point, gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15 (matches cv2 4.13 on every (b, g, r) of a
// 5-step grid and 262144 random triples, tests/test_host_logic.py).  4 pixels per thread: 3 x 32-bit loads, 1 store.
__global__ void __launch_bounds__(256) k_bgr2gray(const uint8_t* __restrict__ bgr, size_t in_pitch_bytes, size_t in_frame_bytes,
                                                  int W, int H, uint8_t* __restrict__ gray, size_t out_pitch_bytes,
                                                  size_t out_frame_bytes) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y, f = blockIdx.z;
    if (x4 >= W || y >= H) return;
    const uint8_t* in = bgr + (size_t)f * in_frame_bytes + (size_t)y * in_pitch_bytes + (size_t)x4 * 3;
    uint8_t* out = gray + (size_t)f * out_frame_bytes + (size_t)y * out_pitch_bytes + x4;
    auto conv = [](unsigned b, unsigned g, unsigned r) { return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15; };
    const bool vec = (x4 + 3 < W) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
    if (vec) {
        const unsigned w0 = __ldg(reinterpret_cast<const unsigned*>(in)), w1 = __ldg(reinterpret_cast<const unsigned*>(in) + 1),
                       w2 = __ldg(reinterpret_cast<const unsigned*>(in) + 2);
        // bytes: b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
        const unsigned p0 = conv(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
        const unsigned p1 = conv(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
        const unsigned p2 = conv((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
        const unsigned p3 = conv((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
        *reinterpret_cast<unsigned*>(out) = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
    } else {
        for (int j = 0; j < 4 && x4 + j < W; ++j) out[j] = (uint8_t)conv(in[3 * j], in[3 * j + 1], in[3 * j + 2]);
    }
}

struct PyrLevelDesc {
    const int* ix; const float* ax; const float* kern;
    float* tmp; size_t tmp_frame_stride;
    int ksize, w, tmp_pitch;
};
constexpr int kPyrMaxLevels = 15;
struct PyrHArgs { PyrLevelDesc lv[kPyrMaxLevels]; int nlev; };

// One CTA per (kPyrRows = 4 source rows, frame).  The four rows are converted to float once and kept INTERLEAVED in
// shared memory -- one float4 per source pixel holding the 4 rows, pixel i in slot i + (i >> 3) -- so a tap is one
// LDS.128 for all rows, conflict-free for the power-of-two lane strides of the decimating levels (lanes 1, 2, 4, 8 source
// pixels apart land in 8 distinct 16-byte bank groups per quarter warp).  Every coarser level takes its horizontally
// blurred + column-interpolated samples from that copy; tap tables are read once per output column.  (The planar layout
// this replaces spent 83 instructions per source pixel, mostly shared-memory index arithmetic; ncu, profiles/.)
constexpr int kPyrRows = 4;
__host__ __device__ inline int pyr_slot(int i) { return i + (i >> 3); }
inline size_t pyr_h_smem_bytes(int W) { return (size_t)(pyr_slot(W) + 1) * sizeof(float4); }

template <typename T>
__global__ void __launch_bounds__(256) k_pyr_h_multi(const T* __restrict__ src, size_t src_pitch_bytes, size_t src_frame_bytes,
                                                     int W, int H, const PyrHArgs pa) {
    extern __shared__ __align__(16) float4 srow4[];
    const int r0 = blockIdx.x * kPyrRows, f = blockIdx.y;
    const char* fbase = (const char*)src + (size_t)f * src_frame_bytes;
    const T* rows[kPyrRows];
#pragma unroll
    for (int rr = 0; rr < kPyrRows; ++rr) rows[rr] = (const T*)(fbase + (size_t)min(r0 + rr, H - 1) * src_pitch_bytes);
    const bool vec = (sizeof(T) == 1) && ((W & 3) == 0) && ((src_pitch_bytes & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(fbase) & 3) == 0);
    if (vec) {
        for (int x = threadIdx.x * 4; x < W; x += 256 * 4) {
            unsigned q[kPyrRows];
#pragma unroll
            for (int rr = 0; rr < kPyrRows; ++rr) q[rr] = __ldg(reinterpret_cast<const unsigned*>(reinterpret_cast<const uint8_t*>(rows[rr]) + x));
#pragma unroll
            for (int b = 0; b < 4; ++b)
                srow4[pyr_slot(x + b)] = make_float4((float)((q[0] >> (8 * b)) & 0xffu), (float)((q[1] >> (8 * b)) & 0xffu),
                                                     (float)((q[2] >> (8 * b)) & 0xffu), (float)((q[3] >> (8 * b)) & 0xffu));
        }
    } else {
        for (int x = threadIdx.x; x < W; x += 256)
            srow4[pyr_slot(x)] = make_float4(load_px(rows[0] + x), load_px(rows[1] + x), load_px(rows[2] + x), load_px(rows[3] + x));
    }
    __syncthreads();
    for (int l = 0; l < pa.nlev; ++l) {
        const PyrLevelDesc& d = pa.lv[l];
        const int rad = d.ksize >> 1, ksize = d.ksize;
        const float* __restrict__ kern = d.kern;
        float* tbase = d.tmp + (size_t)f * d.tmp_frame_stride + (size_t)r0 * d.tmp_pitch;
        for (int x = threadIdx.x; x < d.w; x += 256) {
            const int i0 = d.ix[x];
            const float a = d.ax[x];
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (i0 - rad >= 0 && i0 + 1 + rad < W) {
                int i = i0 - rad;
                float4 prev = srow4[pyr_slot(i)];
                for (int j = 0; j < ksize; ++j) {
                    const float kj = __ldg(kern + j);
                    ++i;
                    const float4 nxt = srow4[pyr_slot(i)];
                    b0.x = fmaf(kj, prev.x, b0.x); b0.y = fmaf(kj, prev.y, b0.y); b0.z = fmaf(kj, prev.z, b0.z); b0.w = fmaf(kj, prev.w, b0.w);
                    b1.x = fmaf(kj, nxt.x, b1.x); b1.y = fmaf(kj, nxt.y, b1.y); b1.z = fmaf(kj, nxt.z, b1.z); b1.w = fmaf(kj, nxt.w, b1.w);
                    prev = nxt;
                }
            } else {
                const int i1 = min(i0 + 1, W - 1);
                for (int j = 0; j < ksize; ++j) {
                    const float kj = __ldg(kern + j);
                    const float4 v0 = srow4[pyr_slot(reflect101(i0 - rad + j, W))], v1 = srow4[pyr_slot(reflect101(i1 - rad + j, W))];
                    b0.x = fmaf(kj, v0.x, b0.x); b0.y = fmaf(kj, v0.y, b0.y); b0.z = fmaf(kj, v0.z, b0.z); b0.w = fmaf(kj, v0.w, b0.w);
                    b1.x = fmaf(kj, v1.x, b1.x); b1.y = fmaf(kj, v1.y, b1.y); b1.z = fmaf(kj, v1.z, b1.z); b1.w = fmaf(kj, v1.w, b1.w);
                }
            }
            const float o[kPyrRows] = {(a != 0.f) ? (b0.x * (1.f - a) + b1.x * a) : b0.x, (a != 0.f) ? (b0.y * (1.f - a) + b1.y * a) : b0.y,
                                       (a != 0.f) ? (b0.z * (1.f - a) + b1.z * a) : b0.z, (a != 0.f) ? (b0.w * (1.f - a) + b1.w * a) : b0.w};
#pragma unroll
            for (int rr = 0; rr < kPyrRows; ++rr)
                if (r0 + rr < H) tbase[(size_t)rr * d.tmp_pitch + x] = o[rr];
        }
    }
}

// Vertical part, 4 adjacent columns per thread (128-bit loads of the horizontal-pass rows, 128-bit store): out[f][y][x] =
// lerp_y(blur_v(tmp)).  Same arithmetic as k_pyr_v (farneback_kernels.cuh), which remains for unaligned pitches.
__global__ void __launch_bounds__(256) k_pyr_v4(const float* __restrict__ tmp, int tmp_pitch, size_t tmp_frame_stride, int H, int w,
                                                int h, const int* __restrict__ iy, const float* __restrict__ ay,
                                                const float* __restrict__ kern, int ksize, float* __restrict__ out, int out_pitch,
                                                size_t out_frame_stride) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (x >= w || y >= h) return;
    const float* t = tmp + (size_t)f * tmp_frame_stride + x;
    const int rad = ksize >> 1;
    const int i0 = iy[y];
    const float a = ay[y];
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    auto ldrow = [&](int r) { return __ldg(reinterpret_cast<const float4*>(t + (size_t)r * tmp_pitch)); };
    if (i0 - rad >= 0 && i0 + 1 + rad < H) {
        const float* q = t + (size_t)(i0 - rad) * tmp_pitch;
        float4 prev = __ldg(reinterpret_cast<const float4*>(q));
        for (int j = 0; j < ksize; ++j) {
            q += tmp_pitch;
            const float4 nxt = __ldg(reinterpret_cast<const float4*>(q));
            const float kj = __ldg(kern + j);
            b0.x = fmaf(kj, prev.x, b0.x); b0.y = fmaf(kj, prev.y, b0.y); b0.z = fmaf(kj, prev.z, b0.z); b0.w = fmaf(kj, prev.w, b0.w);
            b1.x = fmaf(kj, nxt.x, b1.x); b1.y = fmaf(kj, nxt.y, b1.y); b1.z = fmaf(kj, nxt.z, b1.z); b1.w = fmaf(kj, nxt.w, b1.w);
            prev = nxt;
        }
    } else {
        for (int j = 0; j < ksize; ++j) {
            const float kj = __ldg(kern + j);
            const float4 v = ldrow(reflect101(i0 - rad + j, H));
            b0.x = fmaf(kj, v.x, b0.x); b0.y = fmaf(kj, v.y, b0.y); b0.z = fmaf(kj, v.z, b0.z); b0.w = fmaf(kj, v.w, b0.w);
        }
        if (a != 0.f) {
            const int i1 = min(i0 + 1, H - 1);
            for (int j = 0; j < ksize; ++j) {
                const float kj = __ldg(kern + j);
                const float4 v = ldrow(reflect101(i1 - rad + j, H));
                b1.x = fmaf(kj, v.x, b1.x); b1.y = fmaf(kj, v.y, b1.y); b1.z = fmaf(kj, v.z, b1.z); b1.w = fmaf(kj, v.w, b1.w);
            }
        }
    }
    float4 v = b0;
    if (a != 0.f) v = make_float4(b0.x * (1.f - a) + b1.x * a, b0.y * (1.f - a) + b1.y * a, b0.z * (1.f - a) + b1.z * a, b0.w * (1.f - a) + b1.w * a);
    float* op = out + (size_t)f * out_frame_stride + (size_t)y * out_pitch + x;
    if (x + 3 < w) {
        *reinterpret_cast<float4*>(op) = v;
    } else {
        const float vv[4] = {v.x, v.y, v.z, v.w};
        for (int j = 0; j < 4 && x + j < w; ++j) op[j] = vv[j];
    }
}
// pitches and bases that allow the 128-bit version (plan buffers always do: pitch is a multiple of 32 floats)
inline bool pyr_v4_ok(const void* tmp, int tmp_pitch, size_t tmp_frame_stride, const void* out, int out_pitch, size_t out_frame_stride) {
    return aligned16(tmp) && aligned16(out) && (tmp_pitch % 4) == 0 && (out_pitch % 4) == 0 && (tmp_frame_stride % 4) == 0 &&
           (out_frame_stride % 4) == 0;
}

inline bool polyexp_fast_supported(int n, int pitch) { return (n == 5 || n == 7) && (pitch % 4) == 0; }

inline bool polyexp_fast_aligned(const void* R, size_t plane_stride, size_t slot_stride) {
    return aligned16(R) && (plane_stride % 4) == 0 && (slot_stride % 4) == 0;
}

template <int N, bool RH>
inline void launch_polyexp_fast_n(const float* I, int pitch, size_t frame_stride, int w, int h, void* R,
                                  size_t plane_stride, size_t slot_stride, int slot0, int nslots, int nf,
                                  const PolyCoef& pc, cudaStream_t st) {
    using C = FastPeCfg<N>;
    cudaFuncSetAttribute(k_polyexp<N, RH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    dim3 g((w + kFbTW - 1) / kFbTW, (h + kFbTH - 1) / kFbTH, nf);
    k_polyexp<N, RH><<<g, 256, C::SMEM, st>>>(I, pitch, frame_stride, w, h, R, plane_stride, slot_stride, slot0, nslots, pc);
}

template <typename T>
inline void launch_polyexp_l0(const T* src, size_t src_pitch_bytes, size_t src_frame_bytes, int pitch, int w, int h, void* R,
                              size_t plane_stride, size_t slot_stride, int slot0, int nslots, int nf, const PolyCoef& pc,
                              bool r_half, cudaStream_t st) {
    dim3 g((w + kFbTW - 1) / kFbTW, (h + kFbTH - 1) / kFbTH, nf);
#define BF_LAUNCH_L0(NN, RHH)                                                                                             \
    do {                                                                                                                  \
        cudaFuncSetAttribute(k_polyexp_l0<NN, RHH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedPeCfg<NN>::SMEM); \
        k_polyexp_l0<NN, RHH, T><<<g, 256, FusedPeCfg<NN>::SMEM, st>>>(src, src_pitch_bytes, src_frame_bytes, pitch, w, h, R,    \
                                                                       plane_stride, slot_stride, slot0, nslots, pc);      \
    } while (0)
    if (pc.n == 5) { if (r_half) BF_LAUNCH_L0(5, true); else BF_LAUNCH_L0(5, false); }
    else { if (r_half) BF_LAUNCH_L0(7, true); else BF_LAUNCH_L0(7, false); }
#undef BF_LAUNCH_L0
}

inline void launch_polyexp_fast(const float* I, int pitch, size_t frame_stride, int w, int h, void* R,
                                size_t plane_stride, size_t slot_stride, int slot0, int nslots, int nf,
                                const PolyCoef& pc, bool r_half, cudaStream_t st) {
    if (pc.n == 5) {
        if (r_half) launch_polyexp_fast_n<5, true>(I, pitch, frame_stride, w, h, R, plane_stride, slot_stride, slot0, nslots, nf, pc, st);
        else launch_polyexp_fast_n<5, false>(I, pitch, frame_stride, w, h, R, plane_stride, slot_stride, slot0, nslots, nf, pc, st);
    } else {
        if (r_half) launch_polyexp_fast_n<7, true>(I, pitch, frame_stride, w, h, R, plane_stride, slot_stride, slot0, nslots, nf, pc, st);
        else launch_polyexp_fast_n<7, false>(I, pitch, frame_stride, w, h, R, plane_stride, slot_stride, slot0, nslots, nf, pc, st);
    }
}

}  // namespace bf
