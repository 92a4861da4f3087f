// Farneback dense optical flow: generic (runtime-parameter) CUDA kernels for sm_100a.
//
// Algorithm restated from the validated behavioural spec of cv2.calcOpticalFlowFarneback (the single
// hot call of the reference, /root/reference/optical_flow.py:173; spec in SURVEY.md Appendix A).
// Data layout in HBM:
//   level image  I   [frame][h][pitch] f32
//   poly coeffs  R   exact plans:   [slot][5][h][pitch] f32 planes (b_y, b_x, A_yy, A_xx, A_xy)
//                    compact plans: [slot][h][pitch] x 16 B per pixel (RPix: b in fp32, A in fp16), one LDG.128 per bilinear tap
//   matrices     M   [pair][5][h][pitch] planes (G11, G12, G22, h1, h2): f32 (exact) or fp16 (compact)
//   flow             [pair][h][pitch] float2 (dx, dy)
// `pitch` is in elements and a multiple of 32 for plan-owned buffers.  Kernels templated on RH = compact storage.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace bf {

constexpr int kMaxPolyN = 16;
constexpr int kMaxWinHalf = 64;

struct PolyCoef {
    float g[kMaxPolyN + 1];
    float xg[kMaxPolyN + 1];
    float xxg[kMaxPolyN + 1];
    float ig11, ig03, ig33, ig55;
    int n;
};

struct WinCoef {
    float ker[kMaxWinHalf + 1];  // Gaussian window taps ker[0..m] (normalised); unused for box
    float scale;                 // box: 1 / winsize^2 ; Gaussian: 1
    int m;                       // half window
    int gauss;
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

__device__ __forceinline__ float load_px(const uint8_t* p) { return (float)(*p); }
__device__ __forceinline__ float load_px(const float* p) { return *p; }

// ---------------------------------------------------------------------------------------------------
// K1a: horizontal part of (GaussianBlur REFLECT_101 -> bilinear resize), evaluated only at the columns
// the resize samples.  tmp[f][r][x] for every source row r.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_pyr_h(const T* __restrict__ src, size_t src_pitch_bytes, size_t src_frame_bytes, int W, int H,
                        int w, const int* __restrict__ ix, const float* __restrict__ ax,
                        const float* __restrict__ kern, int ksize, float* __restrict__ tmp, int tmp_pitch,
                        size_t tmp_frame_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (x >= w || r >= H) return;
    const T* row = (const T*)((const char*)src + (size_t)f * src_frame_bytes + (size_t)r * src_pitch_bytes);
    const int rad = ksize >> 1;
    const int i0 = ix[x];
    const float a = ax[x];
    float b0 = 0.f, b1 = 0.f;
    if (i0 - rad >= 0 && i0 + 1 + rad < W) {
        // interior: share the overlapping taps of the two neighbouring blurred samples
        float prev = load_px(row + i0 - rad);
        for (int j = 0; j < ksize; ++j) {
            const float nxt = load_px(row + i0 - rad + j + 1);
            const float kj = __ldg(kern + j);
            b0 = fmaf(kj, prev, b0);
            b1 = fmaf(kj, nxt, b1);
            prev = nxt;
        }
    } else {
        for (int j = 0; j < ksize; ++j) {
            const float kj = __ldg(kern + j);
            b0 = fmaf(kj, load_px(row + reflect101(i0 - rad + j, W)), b0);
        }
        if (a != 0.f) {
            const int i1 = min(i0 + 1, W - 1);
            for (int j = 0; j < ksize; ++j)
                b1 = fmaf(__ldg(kern + j), load_px(row + reflect101(i1 - rad + j, W)), b1);
        }
    }
    const float v = (a != 0.f) ? (b0 * (1.f - a) + b1 * a) : b0;
    tmp[(size_t)f * tmp_frame_stride + (size_t)r * tmp_pitch + x] = v;
}

// K1b: vertical part; out[f][y][x] = lerp_y( blur_v(tmp) ).
__global__ void k_pyr_v(const float* __restrict__ tmp, int tmp_pitch, size_t tmp_frame_stride, int H, int w,
                        int h, const int* __restrict__ iy, const float* __restrict__ ay,
                        const float* __restrict__ kern, int ksize, float* __restrict__ out, int out_pitch,
                        size_t out_frame_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (x >= w || y >= h) return;
    const float* t = tmp + (size_t)f * tmp_frame_stride + x;
    const int rad = ksize >> 1;
    const int i0 = iy[y];
    const float a = ay[y];
    float b0 = 0.f, b1 = 0.f;
    if (i0 - rad >= 0 && i0 + 1 + rad < H) {
        float prev = t[(size_t)(i0 - rad) * tmp_pitch];
        for (int j = 0; j < ksize; ++j) {
            const float nxt = t[(size_t)(i0 - rad + j + 1) * tmp_pitch];
            const float kj = __ldg(kern + j);
            b0 = fmaf(kj, prev, b0);
            b1 = fmaf(kj, nxt, b1);
            prev = nxt;
        }
    } else {
        for (int j = 0; j < ksize; ++j)
            b0 = fmaf(__ldg(kern + j), t[(size_t)reflect101(i0 - rad + j, H) * tmp_pitch], b0);
        if (a != 0.f) {
            const int i1 = min(i0 + 1, H - 1);
            for (int j = 0; j < ksize; ++j)
                b1 = fmaf(__ldg(kern + j), t[(size_t)reflect101(i1 - rad + j, H) * tmp_pitch], b1);
        }
    }
    const float v = (a != 0.f) ? (b0 * (1.f - a) + b1 * a) : b0;
    out[(size_t)f * out_frame_stride + (size_t)y * out_pitch + x] = v;
}

// ---------------------------------------------------------------------------------------------------
// K2 (generic): separable polynomial expansion.  Tile 64 x 16, 256 threads.
// Vertical pass straight from global (rows clamped = replicate), horizontal pass from shared memory
// (columns clamped = replicate of the vertical-pass result, exactly as cv2 pads its row buffer).
// ---------------------------------------------------------------------------------------------------
constexpr int kPeTW = 64, kPeTH = 16;

__global__ void __launch_bounds__(256) k_polyexp_generic(const float* __restrict__ I, int pitch,
                                                         size_t frame_stride, int w, int h,
                                                         float* __restrict__ R, size_t plane_stride,
                                                         size_t slot_stride, int slot0, int nslots,
                                                         const PolyCoef pc) {
    __shared__ float s[3][kPeTH][kPeTW + 2 * kMaxPolyN];
    const int n = pc.n;
    const int x0 = blockIdx.x * kPeTW, y0 = blockIdx.y * kPeTH;
    const int f = blockIdx.z;
    const float* img = I + (size_t)f * frame_stride;
    const int tw = kPeTW + 2 * n;
    for (int it = threadIdx.x; it < kPeTH * tw; it += blockDim.x) {
        const int ty = it / tw, tx = it - ty * tw;
        const int gy = y0 + ty;
        if (gy >= h) continue;
        const int gx = min(max(x0 - n + tx, 0), w - 1);
        const float c = img[(size_t)gy * pitch + gx];
        float t0 = c * pc.g[0], t1 = 0.f, t2 = 0.f;
        for (int k = 1; k <= n; ++k) {
            const float up = img[(size_t)max(gy - k, 0) * pitch + gx];
            const float dn = img[(size_t)min(gy + k, h - 1) * pitch + gx];
            const float p = up + dn;
            t0 = fmaf(pc.g[k], p, t0);
            t1 = fmaf(pc.xg[k], dn - up, t1);
            t2 = fmaf(pc.xxg[k], p, t2);
        }
        s[0][ty][tx] = t0;
        s[1][ty][tx] = t1;
        s[2][ty][tx] = t2;
    }
    __syncthreads();
    float* Rb = R + (size_t)((slot0 + f) % nslots) * slot_stride;
    for (int it = threadIdx.x; it < kPeTH * kPeTW; it += blockDim.x) {
        const int ty = it / kPeTW, tx = it - ty * kPeTW;
        const int gx = x0 + tx, gy = y0 + ty;
        if (gx >= w || gy >= h) continue;
        const int cx = tx + n;
        float b1 = s[0][ty][cx] * pc.g[0], b3 = s[1][ty][cx] * pc.g[0], b5 = s[2][ty][cx] * pc.g[0];
        float b2 = 0.f, b4 = 0.f, b6 = 0.f;
        for (int k = 1; k <= n; ++k) {
            const float p0 = s[0][ty][cx + k], m0 = s[0][ty][cx - k];
            const float p1 = s[1][ty][cx + k], m1 = s[1][ty][cx - k];
            const float p2 = s[2][ty][cx + k], m2 = s[2][ty][cx - k];
            const float tg = p0 + m0;
            b1 = fmaf(tg, pc.g[k], b1);
            b4 = fmaf(tg, pc.xxg[k], b4);
            b2 = fmaf(p0 - m0, pc.xg[k], b2);
            b3 = fmaf(p1 + m1, pc.g[k], b3);
            b6 = fmaf(p1 - m1, pc.xg[k], b6);
            b5 = fmaf(p2 + m2, pc.g[k], b5);
        }
        const size_t o = (size_t)gy * pitch + gx;
        Rb[o] = b3 * pc.ig11;
        Rb[plane_stride + o] = b2 * pc.ig11;
        Rb[2 * plane_stride + o] = fmaf(b1, pc.ig03, b5 * pc.ig33);
        Rb[3 * plane_stride + o] = fmaf(b1, pc.ig03, b4 * pc.ig33);
        Rb[4 * plane_stride + o] = b6 * pc.ig55;
    }
}

// ---------------------------------------------------------------------------------------------------
// UpdateMatrices for one pixel: bilinear gather of R1 at (x+dx, y+dy), fallback branch when
// the 2x2 footprint is not strictly inside, border attenuation in the outer 5 px.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float border_w(int i, int n) {
    // table {0.14, 0.14, 0.4472, 0.4472, 0.4472}
    float s = 1.f;
    if (i < 5) s *= (i < 2) ? 0.14f : 0.4472f;
    const int j = n - 1 - i;
    if (j < 5) s *= (j < 2) ? 0.14f : 0.4472f;
    return s;
}

// Addressing: `plane` and `pitch` are 32-bit element counts and every pointer is formed as base + c*plane + offset
// with an unsigned 32-bit offset, which ptxas turns into one IMAD.WIDE per pair of loads (a 64-bit size_t
// formulation costs ~55 integer instructions per pixel here; measured with ncu, see profiles/).
__device__ __forceinline__ void update_px(const float* __restrict__ R0, const float* __restrict__ R1,
                                          unsigned plane, unsigned pitch, int w, int h, int x, int y,
                                          float dx, float dy, float out[5]) {
    float q[5];
    {
        const float* pq = R0 + ((unsigned)y * pitch + (unsigned)x);
#pragma unroll
        for (int c = 0; c < 5; ++c) { q[c] = __ldg(pq); pq += plane; }
    }
    float fx = (float)x + dx, fy = (float)y + dy;
    const int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
    fx -= (float)x1;
    fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float* pa = R1 + ((unsigned)y1 * pitch + (unsigned)x1);
        float t0[5], t1[5], b0[5], b1[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const float* pb = pa + pitch;
            t0[c] = __ldg(pa); t1[c] = __ldg(pa + 1);
            b0[c] = __ldg(pb); b1[c] = __ldg(pb + 1);
            pa += plane;
        }
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        float rw[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) rw[c] = a00 * t0[c] + a01 * t1[c] + a10 * b0[c] + a11 * b1[c];
        r2 = rw[0];
        r3 = rw[1];
        r4 = (q[2] + rw[2]) * 0.5f;
        r5 = (q[3] + rw[3]) * 0.5f;
        r6 = (q[4] + rw[4]) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = q[2];
        r5 = q[3];
        r6 = q[4] * 0.5f;
    }
    r2 = (q[0] - r2) * 0.5f;
    r3 = (q[1] - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float sc = border_w(x, w) * border_w(y, h);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    out[0] = r4 * r4 + r6 * r6;
    out[1] = (r4 + r5) * r6;
    out[2] = r5 * r5 + r6 * r6;
    out[3] = r4 * r2 + r6 * r3;
    out[4] = r6 * r2 + r5 * r3;
}

// Compact plans (RH) also keep the matrices M as fp16 planes: |M| <= ~1.3e3 for uint8 frames (max over the probe set,
// DESIGN.md), far from fp16 overflow; values below fp16's subnormal range vanish against the 1e-3 regulariser.  Measured
// cost with both R and M in fp16: <= 3e-4 px mean, 2.2e-3 px max at 10 px flows (gate: 0.01 / 0.05).  Sums stay fp32.
template <bool RH> struct MStore { using type = typename std::conditional<RH, __half, float>::type; };

__device__ __forceinline__ float m_to_float(float v) { return v; }
__device__ __forceinline__ float m_to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ void m_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void m_from_float(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ float4 m_load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 m_load4(const __half* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    union { unsigned v; __half2 h; } a, b;
    a.v = u.x; b.v = u.y;
    const float2 lo = __half22float2(a.h), hi = __half22float2(b.h);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <typename MT>
__device__ __forceinline__ void store_m(MT* __restrict__ M, unsigned plane, unsigned o, const float m[5]) {
    MT* pm = M + o;
#pragma unroll
    for (int c = 0; c < 5; ++c) { m_from_float(pm, m[c]); pm += plane; }
}

// ---- packed polynomial coefficients: 16 bytes per pixel, one 128-bit load per bilinear tap ------------------------
// Storage format only: all arithmetic stays fp32.  The linear terms b (whose frame-to-frame DIFFERENCE drives the flow)
// stay fp32; the quadratic terms A, which only enter through averages, are fp16: (b_y f32, b_x f32, (A_yy, A_xx) f16x2,
// (A_xy, 0) f16x2).  Against the earlier all-fp16 pixel this unpacks with 3 conversions per tap instead of 5 and removes
// most of the storage error (tests/test_gpu_flow.py).  It turns the 20 scalar gather loads + 5 centre
// loads per pixel of the planar layout into 4 + 1 LDG.128.  Used for uint8 input only (|A| is bounded by the 0..255 range).
struct __align__(16) RPix { float by, bx; __half2 ayy_axx, axy_0; };

__device__ __forceinline__ uint4 pack_r(float r0, float r1, float r2, float r3, float r4) {
    union { unsigned v; __half2 h; } a, b;
    a.h = __floats2half2_rn(r2, r3);
    b.h = __floats2half2_rn(r4, 0.f);
    return make_uint4(__float_as_uint(r0), __float_as_uint(r1), a.v, b.v);
}

__device__ __forceinline__ void unpack_r(const uint4& u, float v[5]) {
    union { unsigned v; __half2 h; } a, b;
    a.v = u.z; b.v = u.w;
    const float2 q = __half22float2(a.h);
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = q.x; v[3] = q.y; v[4] = __low2float(b.h);
}

// UpdateMatrices for one pixel from packed R (same arithmetic as update_px).  R0/R1 point at pixel (0,0) of the frame.
template <bool BORDER = true>
__device__ __forceinline__ void update_px_h(const uint4* __restrict__ R0, const uint4* __restrict__ R1, unsigned pitch,
                                            int w, int h, int x, int y, float dx, float dy, float out[5]) {
    float q[5];
    unpack_r(__ldg(R0 + (unsigned)y * pitch + (unsigned)x), q);
    float fx = (float)x + dx, fy = (float)y + dy;
    const int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
    fx -= (float)x1;
    fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const uint4* pa = R1 + (unsigned)y1 * pitch + (unsigned)x1;
        const uint4 u00 = __ldg(pa), u01 = __ldg(pa + 1), u10 = __ldg(pa + pitch), u11 = __ldg(pa + pitch + 1);
        float t00[5], t01[5], t10[5], t11[5];
        unpack_r(u00, t00); unpack_r(u01, t01); unpack_r(u10, t10); unpack_r(u11, t11);
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        float rw[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) rw[c] = a00 * t00[c] + a01 * t01[c] + a10 * t10[c] + a11 * t11[c];
        r2 = rw[0];
        r3 = rw[1];
        r4 = (q[2] + rw[2]) * 0.5f;
        r5 = (q[3] + rw[3]) * 0.5f;
        r6 = (q[4] + rw[4]) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = q[2];
        r5 = q[3];
        r6 = q[4] * 0.5f;
    }
    r2 = (q[0] - r2) * 0.5f;
    r3 = (q[1] - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if (BORDER && ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10))) {
        const float sc = border_w(x, w) * border_w(y, h);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    out[0] = r4 * r4 + r6 * r6;
    out[1] = (r4 + r5) * r6;
    out[2] = r5 * r5 + r6 * r6;
    out[3] = r4 * r2 + r6 * r3;
    out[4] = r6 * r2 + r5 * r3;
}

// The same update split in two so that the gather of the NEXT pixel is in flight while this one is being computed (the
// one-piece version issues its four taps inside the `inside` branch, so every pixel pays a full L2 round trip in series;
// ncu source page, profiles/).  Branch-free: the taps are always fetched from a clamped footprint (requires w, h >= 2)
// and the fallback branch becomes five selects.  Arithmetic and operation order are those of update_px_h.
struct UpdTaps { uint4 q, u00, u01, u10, u11; float fx, fy, dx, dy; bool inside; };

__device__ __forceinline__ void update_issue_h(const uint4* __restrict__ R0, const uint4* __restrict__ R1, unsigned pitch,
                                               int w, int h, int x, int y, float dx, float dy, UpdTaps& t) {
    t.q = __ldg(R0 + (unsigned)y * pitch + (unsigned)x);
    const float fx = (float)x + dx, fy = (float)y + dy;
    const int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
    t.fx = fx - (float)x1;
    t.fy = fy - (float)y1;
    t.dx = dx; t.dy = dy;
    t.inside = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
    const int cx = min(max(x1, 0), w - 2), cy = min(max(y1, 0), h - 2);
    const uint4* pa = R1 + (unsigned)cy * pitch + (unsigned)cx;
    t.u00 = __ldg(pa); t.u01 = __ldg(pa + 1); t.u10 = __ldg(pa + pitch); t.u11 = __ldg(pa + pitch + 1);
}

template <bool BORDER>
__device__ __forceinline__ void update_finish_h(const UpdTaps& t, int w, int h, int x, int y, float out[5]) {
    float q[5], t00[5], t01[5], t10[5], t11[5];
    unpack_r(t.q, q);
    unpack_r(t.u00, t00); unpack_r(t.u01, t01); unpack_r(t.u10, t10); unpack_r(t.u11, t11);
    const float fx = t.fx, fy = t.fy;
    const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
    float rw[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) rw[c] = a00 * t00[c] + a01 * t01[c] + a10 * t10[c] + a11 * t11[c];
    const bool in = t.inside;
    float r2 = in ? rw[0] : 0.f;
    float r3 = in ? rw[1] : 0.f;
    float r4 = in ? (q[2] + rw[2]) * 0.5f : q[2];
    float r5 = in ? (q[3] + rw[3]) * 0.5f : q[3];
    float r6 = in ? (q[4] + rw[4]) * 0.25f : q[4] * 0.5f;
    r2 = (q[0] - r2) * 0.5f;
    r3 = (q[1] - r3) * 0.5f;
    r2 += r4 * t.dy + r6 * t.dx;
    r3 += r6 * t.dy + r5 * t.dx;
    if (BORDER && ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10))) {
        const float sc = border_w(x, w) * border_w(y, h);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    out[0] = r4 * r4 + r6 * r6;
    out[1] = (r4 + r5) * r6;
    out[2] = r5 * r5 + r6 * r6;
    out[3] = r4 * r2 + r6 * r3;
    out[4] = r6 * r2 + r5 * r3;
}

// R0/R1 of pair p in the frame ring (p < nslots, slot0 < nslots: one conditional subtract instead of an integer modulo).
__device__ __forceinline__ int ring_slot(int slot0, int p, int nslots) {
    const int s = slot0 + p;
    return s >= nslots ? s - nslots : s;
}

// Layout-agnostic front end: RH = packed fp16 (slot_stride counts uint4 pixels), else fp32 planes (slot_stride in floats).
// BORDER = false: the caller guarantees the pixel lies outside the 5-px attenuation ring (interior tiles).
template <bool RH, bool BORDER = true>
__device__ __forceinline__ void update_px_any(const void* R0, const void* R1, unsigned plane, unsigned pitch, int w, int h,
                                              int x, int y, float dx, float dy, float out[5]) {
    if (RH) update_px_h<BORDER>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), pitch, w, h, x, y, dx, dy, out);
    else update_px(static_cast<const float*>(R0), static_cast<const float*>(R1), plane, pitch, w, h, x, y, dx, dy, out);
}

template <bool RH>
__device__ __forceinline__ const void* r_slot_ptr(const void* R, size_t slot_stride, int slot) {
    if (RH) return static_cast<const uint4*>(R) + (size_t)slot * slot_stride;
    return static_cast<const float*>(R) + (size_t)slot * slot_stride;
}

struct ResizeTab {
    const int* ix; const float* ax;  // [w]
    const int* iy; const float* ay;  // [h]
};

// K3a: M = UpdateMatrices(R0, R1, flow_init).  flow_mode: 0 = zero, 1 = flow buffer, 2 = upsample coarse.
struct UpdateArgs {
    const void* R; size_t plane_stride, slot_stride; int slot0, nslots;   // R0 = ring slot (slot0+p), R1 = the next one
    int pitch, w, h;
    int flow_mode;
    const float2* flow; int flow_pitch; size_t flow_stride;              // mode 1: [pair][h][flow_pitch]; mode 2: coarse
    int ws, hs; float mult; ResizeTab tab;
    void* M; size_t m_stride;                                              // [pair][5][h][pitch] f32, or fp16 when RH
    float2* flow_out; int flow_out_pitch; size_t flow_out_stride;          // optional: write flow_init
};

// Each thread handles kUpdRows rows of one column: the x-side resize table entry, the ring-slot pointers and the kernel
// parameters are fetched once per thread instead of once per pixel (the per-pixel version was ~330 SASS instructions, a
// third of them index arithmetic and constant loads; ncu, profiles/).
constexpr int kUpdRows = 4;   // 8 rows: 74 registers, slower (18.1 vs 17.3 ms per 128 pairs)

template <bool RH>
__global__ void __launch_bounds__(256) k_update(const UpdateArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yb = (blockIdx.y * blockDim.y + threadIdx.y) * kUpdRows;
    const int p = blockIdx.z;
    if (x >= a.w || yb >= a.h) return;
    using MT = typename MStore<RH>::type;
    const int w = a.w, h = a.h;
    const unsigned pitch = (unsigned)a.pitch, plane = (unsigned)a.plane_stride;
    const void* R0 = nullptr;
    const void* R1 = nullptr;
    MT* Mo = nullptr;
    if (a.M) {
        R0 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p, a.nslots));
        R1 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p + 1, a.nslots));
        Mo = static_cast<MT*>(a.M) + (size_t)p * a.m_stride;
    }
    const float2* fin = a.flow ? a.flow + (size_t)p * a.flow_stride : nullptr;
    float2* fout = a.flow_out ? a.flow_out + (size_t)p * a.flow_out_stride : nullptr;
    int ux0 = 0, ux1 = 0;
    float ua = 0.f;
    if (a.flow_mode == 2) { ux0 = a.tab.ix[x]; ua = a.tab.ax[x]; ux1 = min(ux0 + 1, a.ws - 1); }
    // a block of rows away from the 5-px ring needs no attenuation test (block-uniform)
    const bool inner = (x >= 5) && (x < w - 5) && (yb >= 5) && (yb + kUpdRows <= h - 5);
    // flow_init of the thread's rows first (independent loads), then the updates
    float2 fl[kUpdRows];
    if (a.flow_mode == 2) {
        // bilinear sample of the coarser flow (cv2.resize INTER_LINEAR) times mult.  The rows of a warp are
        // uniform, and consecutive fine rows share coarse rows (4 fine rows touch 4 coarse rows at pyr_scale 0.5, not 8):
        // the x-interpolated coarse rows are carried from one fine row to the next.
        const float ub = 1.f - ua;
        auto hrow = [&](int yc) -> float2 {
            const unsigned r = (unsigned)yc * (unsigned)a.flow_pitch;
            const float2 p0 = fin[r + (unsigned)ux0], p1 = fin[r + (unsigned)ux1];
            return make_float2(p0.x * ub + p1.x * ua, p0.y * ub + p1.y * ua);
        };
        int c0 = -1, c1 = -1;
        float2 H0 = make_float2(0.f, 0.f), H1 = H0;
#pragma unroll
        for (int k = 0; k < kUpdRows; ++k) {
            const int y = min(yb + k, h - 1);
            const int y0 = a.tab.iy[y], y1 = min(y0 + 1, a.hs - 1);
            const float b = a.tab.ay[y];
            const float2 h0 = (y0 == c0) ? H0 : ((y0 == c1) ? H1 : hrow(y0));
            const float2 h1 = (y1 == y0) ? h0 : ((y1 == c1) ? H1 : ((y1 == c0) ? H0 : hrow(y1)));
            c0 = y0; H0 = h0; c1 = y1; H1 = h1;
            fl[k].x = (h0.x * (1.f - b) + h1.x * b) * a.mult;
            fl[k].y = (h0.y * (1.f - b) + h1.y * b) * a.mult;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kUpdRows; ++k) {
            const int y = min(yb + k, h - 1);
            fl[k] = a.flow_mode == 1 ? fin[(unsigned)y * (unsigned)a.flow_pitch + (unsigned)x] : make_float2(0.f, 0.f);
        }
    }
    if (fout) {
#pragma unroll
        for (int k = 0; k < kUpdRows; ++k)
            if (yb + k < h) fout[(unsigned)(yb + k) * (unsigned)a.flow_out_pitch + (unsigned)x] = fl[k];
    }
    if (!Mo) return;
    if (RH && w >= 2 && h >= 2) {
        // packed coefficients: row k+1's taps are in flight while row k is computed (update_issue_h / update_finish_h)
        const uint4* R0h = static_cast<const uint4*>(R0);
        const uint4* R1h = static_cast<const uint4*>(R1);
        UpdTaps A, B;
        update_issue_h(R0h, R1h, pitch, w, h, x, min(yb, h - 1), fl[0].x, fl[0].y, A);
#pragma unroll
        for (int k = 0; k < kUpdRows; k += 2) {
            update_issue_h(R0h, R1h, pitch, w, h, x, min(yb + k + 1, h - 1), fl[k + 1].x, fl[k + 1].y, B);
            float m[5];
            if (inner) update_finish_h<false>(A, w, h, x, yb + k, m); else update_finish_h<true>(A, w, h, x, yb + k, m);
            if (yb + k < h) store_m(Mo, plane, (unsigned)(yb + k) * pitch + (unsigned)x, m);
            if (k + 2 < kUpdRows) update_issue_h(R0h, R1h, pitch, w, h, x, min(yb + k + 2, h - 1), fl[k + 2].x, fl[k + 2].y, A);
            if (inner) update_finish_h<false>(B, w, h, x, yb + k + 1, m); else update_finish_h<true>(B, w, h, x, yb + k + 1, m);
            if (yb + k + 1 < h) store_m(Mo, plane, (unsigned)(yb + k + 1) * pitch + (unsigned)x, m);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < kUpdRows; ++k) {
        const int y = yb + k;
        if (y >= h) break;
        float m[5];
        if (inner) update_px_any<RH, false>(R0, R1, plane, pitch, w, h, x, y, fl[k].x, fl[k].y, m);
        else update_px_any<RH, true>(R0, R1, plane, pitch, w, h, x, y, fl[k].x, fl[k].y, m);
        store_m(Mo, plane, (unsigned)y * pitch + (unsigned)x, m);
    }
}

// ---------------------------------------------------------------------------------------------------
// K3b (generic): flow = Solve(Blur(M)) [+ M' = UpdateMatrices(flow)] [+ projection and ROI sums].
// Tile 32 x 8 outputs, 256 threads.  Phase 1: vertical window sums global -> shared (one thread per
// (channel, column), exact first window then short-history sliding: no long-range cancellation).
// Phase 2: horizontal sums from shared, accurate 2x2 solve.  Phase 3: fused tail.
// ---------------------------------------------------------------------------------------------------
constexpr int kBsTW = 32, kBsTH = 8;
constexpr int kRoiVals = 4;  // sum vx, sum vy, sum mag, count

struct BlurSolveArgs {
    const void* M; size_t m_stride, plane_stride; int pitch, w, h;         // f32 planes, or fp16 planes when RH
    // outputs (each optional)
    float2* flow; int flow_pitch; size_t flow_stride;
    void* Mout;
    const void* R; size_t slot_stride; int slot0, nslots;   // for Mout (fp32 planes, or packed fp16 pixels)
    // ROI reduction (optional): masks [n_roi][h][w] u8; axes per pair; partial [pair][roi][ncta][4]
    const uint8_t* masks; int n_roi; size_t mask_stride; int mask_pitch;
    const float* axes;  // [pair][4] = ex0, ex1, ey0, ey1
    float* partial;
};

// accurate a*b - c*d (Kahan): the structure-tensor determinant cancels heavily where the window holds
// 1-D structure; cv2 does this solve in double.
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d) {
    const float cd = c * d;
    const float err = fmaf(-c, d, cd);
    const float dop = fmaf(a, b, -cd);
    return dop + err;
}

__device__ __forceinline__ float2 solve2x2(float g11, float g12, float g22, float h1, float h2) {
    const float det = diff_of_products(g11, g22, g12, g12) + 1e-3f;
    const float idet = 1.f / det;
    float2 r;
    r.x = diff_of_products(g11, h2, g12, h1) * idet;
    r.y = diff_of_products(g22, h1, g12, h2) * idet;
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide ROI partial sums for one pixel per thread; writes partial[...][4] per CTA (deterministic).
__device__ __forceinline__ void roi_reduce_store(const BlurSolveArgs& a, int p, int x, int y, bool valid,
                                                 float2 fl, float* s_red /*[8][4]*/) {
    const float* ax = a.axes + p * 4;
    const float vx = fl.x * ax[0] + fl.y * ax[1];
    const float vy = fl.x * ax[2] + fl.y * ax[3];
    const float mg = sqrtf(vx * vx + vy * vy);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarp = blockDim.x >> 5;
    const int ncta = gridDim.x * gridDim.y;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    for (int r = 0; r < a.n_roi; ++r) {
        const bool in = valid && a.masks[(size_t)r * a.mask_stride + (size_t)y * a.mask_pitch + x] != 0;
        const float s0 = warp_sum(in ? vx : 0.f), s1 = warp_sum(in ? vy : 0.f), s2 = warp_sum(in ? mg : 0.f),
                    s3 = warp_sum(in ? 1.f : 0.f);
        __syncthreads();
        if (lane == 0) {
            s_red[warp * 4 + 0] = s0; s_red[warp * 4 + 1] = s1; s_red[warp * 4 + 2] = s2; s_red[warp * 4 + 3] = s3;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            float t = 0.f;
            for (int i = 0; i < nwarp; ++i) t += s_red[i * 4 + threadIdx.x];
            a.partial[(((size_t)p * a.n_roi + r) * ncta + cta) * kRoiVals + threadIdx.x] = t;
        }
    }
}

template <bool RH>
__global__ void __launch_bounds__(256) k_blur_solve_generic(const BlurSolveArgs a, const WinCoef wc) {
    __shared__ float V[5][kBsTH][kBsTW + 2 * kMaxWinHalf];
    __shared__ float s_red[8 * 4];
    const int m = wc.m;
    const int x0 = blockIdx.x * kBsTW, y0 = blockIdx.y * kBsTH;
    const int p = blockIdx.z;
    using MT = typename MStore<RH>::type;
    const MT* Mp = static_cast<const MT*>(a.M) + (size_t)p * a.m_stride;
    const int tw = kBsTW + 2 * m;
    const int w = a.w, h = a.h;
    for (int it = threadIdx.x; it < 5 * tw; it += blockDim.x) {
        const int c = it / tw, tx = it - c * tw;
        const int gx = min(max(x0 - m + tx, 0), w - 1);
        const MT* col = Mp + c * a.plane_stride + gx;
        auto at = [&](int row) { return m_to_float(col[(size_t)row * a.pitch]); };
        if (wc.gauss) {
            for (int ty = 0; ty < kBsTH; ++ty) {
                const int gy = min(y0 + ty, h - 1);
                float s = at(gy) * wc.ker[0];
                for (int i = 1; i <= m; ++i) s += (at(max(gy - i, 0)) + at(min(gy + i, h - 1))) * wc.ker[i];
                V[c][ty][tx] = s;
            }
        } else {
            float s = 0.f;
            for (int i = -m; i <= m; ++i) s += at(min(max(y0 + i, 0), h - 1));
            V[c][0][tx] = s;
            for (int ty = 1; ty < kBsTH; ++ty) {
                const int gy = y0 + ty;
                s += at(min(gy + m, h - 1)) - at(min(max(gy - m - 1, 0), h - 1));
                V[c][ty][tx] = s;
            }
        }
    }
    __syncthreads();
    const int tx = threadIdx.x & (kBsTW - 1), ty = threadIdx.x / kBsTW;
    const int x = x0 + tx, y = y0 + ty;
    const bool valid = (x < w) && (y < h);
    float g[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const float* v = &V[c][ty][tx + m];
        float s;
        if (wc.gauss) {
            s = v[0] * wc.ker[0];
            for (int i = 1; i <= m; ++i) s += (v[-i] + v[i]) * wc.ker[i];
        } else {
            s = v[0];
            for (int i = 1; i <= m; ++i) s += v[-i] + v[i];
            s *= wc.scale;
        }
        g[c] = s;
    }
    const float2 fl = solve2x2(g[0], g[1], g[2], g[3], g[4]);
    if (valid) {
        if (a.flow) a.flow[(size_t)p * a.flow_stride + (size_t)y * a.flow_pitch + x] = fl;
        if (a.Mout) {
            const void* R0 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p, a.nslots));
            const void* R1 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p + 1, a.nslots));
            float mm[5];
            update_px_any<RH>(R0, R1, (unsigned)a.plane_stride, (unsigned)a.pitch, w, h, x, y, fl.x, fl.y, mm);
            store_m(static_cast<MT*>(a.Mout) + (size_t)p * a.m_stride, (unsigned)a.plane_stride,
                    (unsigned)y * (unsigned)a.pitch + (unsigned)x, mm);
        }
    }
    if (a.partial) roi_reduce_store(a, p, x, y, valid, fl, s_red);
}

// Projection + ROI partial sums of an existing flow buffer (iterations == 0 path; same partial layout).
__global__ void __launch_bounds__(256) k_roi_from_flow(const BlurSolveArgs a) {
    __shared__ float s_red[8 * 4];
    const int tx = threadIdx.x & (kBsTW - 1), ty = threadIdx.x / kBsTW;
    const int x = blockIdx.x * kBsTW + tx, y = blockIdx.y * kBsTH + ty;
    const int p = blockIdx.z;
    const bool valid = (x < a.w) && (y < a.h);
    float2 fl = make_float2(0.f, 0.f);
    if (valid) fl = a.flow[(size_t)p * a.flow_stride + (size_t)y * a.flow_pitch + x];
    roi_reduce_store(a, p, x, y, valid, fl, s_red);
}

// Finalise the ROI means: out[roi][t_first+p][3] = sums / count.  One warp per (pair, roi); lanes stride over the
// per-CTA partials and accumulate in double, then a fixed-order shuffle tree: deterministic run to run.
__global__ void k_roi_finalize(const float* __restrict__ partial, int n_pairs, int n_roi, int ncta,
                               const double* __restrict__ ex, const double* __restrict__ ey, int t_first,
                               float* __restrict__ out, int T) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n_pairs * n_roi) return;
    const int p = i / n_roi, r = i - p * n_roi;
    const float4* q = reinterpret_cast<const float4*>(partial + (size_t)i * ncta * kRoiVals);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int c = lane; c < ncta; c += 32) {
        const float4 v = q[c];
        s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if (lane != 0) return;
    const int t = t_first + p;
    float* o = out + ((size_t)r * T + t) * 3;
    const bool ok = isfinite(ex[2 * t]) && isfinite(ex[2 * t + 1]) && isfinite(ey[2 * t]) && isfinite(ey[2 * t + 1]);
    if (!ok || s3 == 0.0) {
        const float nanv = __int_as_float(0x7fc00000);
        o[0] = o[1] = o[2] = nanv;
    } else {
        o[0] = (float)(s0 / s3); o[1] = (float)(s1 / s3); o[2] = (float)(s2 / s3);
    }
}

// axes[p] = (float)ex[t], ... for pairs t = t_first + p (python float -> float32 as numpy does, optical_flow.py:180-181)
__global__ void k_axes_to_f32(const double* __restrict__ ex, const double* __restrict__ ey, int t_first, int n_pairs,
                              float* __restrict__ axes) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int t = t_first + p;
    axes[p * 4 + 0] = (float)ex[2 * t];
    axes[p * 4 + 1] = (float)ex[2 * t + 1];
    axes[p * 4 + 2] = (float)ey[2 * t];
    axes[p * 4 + 3] = (float)ey[2 * t + 1];
}

__global__ void k_fill_nan(float* __restrict__ p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = __int_as_float(0x7fc00000);
}

}  // namespace bf
