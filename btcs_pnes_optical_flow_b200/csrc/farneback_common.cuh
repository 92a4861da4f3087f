// Farneback dense optical flow: shared device code (layouts, UpdateMatrices arithmetic, ROI reduction) for sm_100a.
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
template <bool BORDER = true>
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

// ---- matrices M of one pair ------------------------------------------------------------------------------------------
// Exact plans: five fp32 planes (G11, G12, G22, h1, h2), 20 B per pixel, rows of `pitch` elements.
// Compact plans (RH): 14 B per pixel -- G11, G12, G22 as fp16 and h1, h2 as fp32 -- with CONSISTENT rounding: the
// structure-tensor terms G are rounded to fp16 first and h = (A b) + Gq d is then formed in fp32 from the ROUNDED G.  The
// blurred system sum(Gq_i) d = sum(h_i) is a weighted mean of the per-pixel solutions, so perturbing the weights G_i by
// 2^-12 moves the result only in proportion to the SPREAD of the flow inside the window, not to its magnitude: measured
// against cv2 (NumPy emulation, tools/emulate_storage.py) this storage gives 4e-7 .. 2e-6 px mean / <= 2e-4 px interior
// max, where all-fp16 matrices gave 1e-4 / 3e-3 (the error came from rounding h, which has to carry G d to full
// precision).  G = A^2 terms stay far inside the fp16 range for uint8 frames (|A| <= ~48); the conversion saturates
// (satfinite) so that an out-of-range value can never turn into inf/NaN in the window sums.
//
// Compact matrices are stored in BLOCKS of 128 columns x 16 rows (one tile of the iteration kernel), 28 KB each:
//     block (bx, by) at ((by * nbx) + bx) * kMbBytes;  inside it  G_c at c * 4096 + (row * 128 + col) * 2   (c = 0..2)
//                                                                 h_c at 12288 + c * 8192 + (row * 128 + col) * 4   (c = 0..1)
// so that, relative to one per-thread pointer, every row and channel a thread touches sits at a COMPILE-TIME offset: the
// vertical window walk of phase 1 and the five stores of the update tail need no address arithmetic per row (the planar
// layout cost ~5 integer instructions per load and ~20 per pixel of stores; ncu, profiles/r2b), and a tile's own block is one
// contiguous 28 KB span (one bulk L2 prefetch).  Padding rows/columns of edge blocks are never written and stay zero.
constexpr int kMbW = 128, kMbH = 16;
constexpr unsigned kMbGBytes = kMbW * kMbH * 2, kMbHBytes = kMbW * kMbH * 4;          // one channel of a block
constexpr unsigned kMbHOff = 3 * kMbGBytes;
constexpr unsigned kMbBytes = 3 * kMbGBytes + 2 * kMbHBytes;                          // 28672
__host__ __device__ inline size_t m_pair_bytes(bool compact, int w, int h, size_t plane) {
    return compact ? (size_t)((w + kMbW - 1) / kMbW) * ((h + kMbH - 1) / kMbH) * kMbBytes : plane * 20u;
}

// One pixel of M on its way to memory.  Exact: five floats.  Compact: the three G terms already rounded to fp16 (raw
// halves, two packed in g01) and the two fp32 h terms formed from those rounded values.
template <bool RH> struct MOut;
template <> struct MOut<false> { float m[5]; };
template <> struct MOut<true> { unsigned g01; unsigned short g2; float h1, h2; };

template <bool RH> struct MView;
template <> struct MView<false> {
    float* p; unsigned plane, pitch;
    __device__ __forceinline__ MView(void* base, size_t pair_stride_bytes, int pair, unsigned plane_, unsigned pitch_, int)
        : p(reinterpret_cast<float*>(static_cast<char*>(base) + (size_t)pair * pair_stride_bytes)), plane(plane_), pitch(pitch_) {}
    __device__ __forceinline__ float load(int c, int y, int x) const { return __ldg(p + (size_t)c * plane + (unsigned)y * pitch + (unsigned)x); }
    __device__ __forceinline__ void store(int y, int x, const MOut<false>& v) const {
        float* pm = p + (unsigned)y * pitch + (unsigned)x;
#pragma unroll
        for (int c = 0; c < 5; ++c) { *pm = v.m[c]; pm += plane; }
    }
};
template <> struct MView<true> {
    char* base; unsigned nbx;
    __device__ __forceinline__ MView(void* base_, size_t pair_stride_bytes, int pair, unsigned, unsigned, int w)
        : base(static_cast<char*>(base_) + (size_t)pair * pair_stride_bytes), nbx((unsigned)(w + kMbW - 1) / kMbW) {}
    __device__ __forceinline__ char* block(int bx, int by) const { return base + ((size_t)((unsigned)by * nbx + (unsigned)bx)) * kMbBytes; }
    __device__ __forceinline__ unsigned block_row_bytes() const { return nbx * kMbBytes; }      // from block (bx, by) to (bx, by + 1)
    // byte offset of pixel (row, col) of G channel 0 / h channel 0 inside a block
    static __device__ __forceinline__ unsigned g_off(int row, int col) { return (unsigned)(row * kMbW + col) * 2u; }
    static __device__ __forceinline__ unsigned h_off(int row, int col) { return kMbHOff + (unsigned)(row * kMbW + col) * 4u; }
    __device__ __forceinline__ float load(int c, int y, int x) const {
        const char* b = block(x >> 7, y >> 4);
        const int row = y & 15, col = x & 127;
        if (c < 3) return __half2float(__ldg(reinterpret_cast<const __half*>(b + c * kMbGBytes + g_off(row, col))));
        return __ldg(reinterpret_cast<const float*>(b + (c - 3) * kMbHBytes + h_off(row, col)));
    }
    // stores through pointers already positioned on the pixel: pg inside G channel 0, ph inside h channel 0 of the block
    static __device__ __forceinline__ void store_at(char* pg, char* ph, const MOut<true>& v) {
        *reinterpret_cast<unsigned short*>(pg) = (unsigned short)(v.g01 & 0xffffu);
        *reinterpret_cast<unsigned short*>(pg + kMbGBytes) = (unsigned short)(v.g01 >> 16);
        *reinterpret_cast<unsigned short*>(pg + 2 * kMbGBytes) = v.g2;
        *reinterpret_cast<float*>(ph) = v.h1;
        *reinterpret_cast<float*>(ph + kMbHBytes) = v.h2;
    }
    __device__ __forceinline__ void store(int y, int x, const MOut<true>& v) const {
        char* b = block(x >> 7, y >> 4);
        store_at(b + g_off(y & 15, x & 127), b + h_off(y & 15, x & 127), v);
    }
};

// ---- mixed-precision helpers (sm_100a: f32 <- f16 (x f16) + f32 in one instruction, exact conversions included) --------
__device__ __forceinline__ float fh_add(unsigned short h, float s) { float d; asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(s)); return d; }
__device__ __forceinline__ float fh_sub(unsigned short h, float s) { float d; asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(s)); return d; }   // h - s
__device__ __forceinline__ float fh_fma(unsigned short a, unsigned short b, float c) { float d; asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c)); return d; }
__device__ __forceinline__ void split_h2(unsigned v, unsigned short& lo, unsigned short& hi) { asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(v)); }
// {lo, hi} -> packed fp16 pair, round to nearest even, saturating to the largest finite value
__device__ __forceinline__ unsigned pack_h2_sat(float lo, float hi) { unsigned d; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo)); return d; }
__device__ __forceinline__ unsigned short f2h_sat(float v) { unsigned short d; asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(d) : "f"(v)); return d; }
__device__ __forceinline__ float h2f(unsigned short h) { float d; asm("cvt.f32.f16 %0, %1;" : "=f"(d) : "h"(h)); return d; }

// ---- packed fp32 pairs (sm_100a fma/add/mul.rn.f32x2 -> FFMA2 / FADD2 / FMUL2: two IEEE fp32 operations per instruction; a scalar
// operand is broadcast when both halves of a pair are the same register).  Measured (tools/ubench_f32x2.cu): the same lane rate as
// scalar FFMA (128 per clock and SM) AND the same issue cost -- 8 FFMA2 + 16 IADD take as long as 16 FFMA + 16 IADD -- so a
// packed instruction buys nothing by itself.  What it gives the kernels that use it is the pair-wide task: one load, one
// conversion, one address and one loop step feed two accumulators, and the code stays half as long. ----------------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 up2(f32x2_t v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// ---- packed polynomial coefficients: 16 bytes per pixel, one 128-bit load per bilinear tap ------------------------
// Storage format only.  The linear terms b (whose frame-to-frame DIFFERENCE drives the flow) stay fp32; the quadratic
// terms A, which only enter through averages, are fp16: (b_y f32, b_x f32, (A_yy, A_xx) f16x2, (A_xy, 0) f16x2).  It turns
// the 20 scalar gather loads + 5 centre loads per pixel of the planar layout into 4 + 1 LDG.128, and the A terms are
// consumed as they are by mixed-precision FMAs (FHFMA: f16 x f16 + f32 -> f32, products exact), so no unpacking
// conversions are issued.  Used for uint8 input only (|A| is bounded by the 0..255 range).  Cost against cv2 (emulation
// and GPU tests): ~1e-6 px mean, <= 2e-5 px interior max.
struct __align__(16) RPix { float by, bx; __half2 ayy_axx, axy_0; };

__device__ __forceinline__ uint4 pack_r(float r0, float r1, float r2, float r3, float r4) {
    union { unsigned v; __half2 h; } a, b;
    a.h = __floats2half2_rn(r2, r3);
    b.h = __floats2half2_rn(r4, 0.f);
    return make_uint4(__float_as_uint(r0), __float_as_uint(r1), a.v, b.v);
}

// UpdateMatrices from packed R, split in two so that the gather of the NEXT pixel can be in flight while this one is being
// computed.  Branch-free: the four taps always come from a clamped footprint, and when the footprint is not strictly inside
// (the reference's fallback branch: A from R0 alone, b1 := 0) the interpolation weights are zeroed -- already at issue
// time -- which yields exactly the fallback values from the same arithmetic.
struct UpdTaps { uint4 q, u00, u01, u10, u11; float fx, gx, fy, sa, dx, dy; };

// r0px: R0 at pixel (x, y) (callers walking a column carry it as a pointer); R1: pixel (0, 0) of the next frame.
// TINY: the image may be one pixel wide or high (runtime-parameter kernel); otherwise w, h >= 2 and the +1 taps sit at
// immediate offsets.
template <bool TINY = false>
__device__ __forceinline__ void update_issue_h(const uint4* __restrict__ r0px, const uint4* __restrict__ R1, unsigned pitch,
                                               int w, int h, int x, int y, float dx, float dy, UpdTaps& t) {
    t.q = __ldg(r0px);
    const float px = (float)x + dx, py = (float)y + dy;
    const int x1 = __float2int_rd(px), y1 = __float2int_rd(py);
    const float fx = px - (float)x1;
    const bool in = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
    t.fx = in ? fx : 0.f;
    t.gx = in ? 1.f - fx : 0.f;
    t.fy = py - (float)y1;
    t.sa = in ? 0.5f : 1.f;
    t.dx = dx; t.dy = dy;
    const int cx = max(min(x1, w - 2), 0), cy = max(min(y1, h - 2), 0);
    const uint4* pa = R1 + (unsigned)cy * pitch + (unsigned)cx;
    if (TINY) {
        const int ox = (w > 1) ? 1 : 0;
        const unsigned oy = (h > 1) ? pitch : 0u;
        t.u00 = __ldg(pa); t.u01 = __ldg(pa + ox); t.u10 = __ldg(pa + oy); t.u11 = __ldg(pa + oy + ox);
    } else {
        const uint4* pb = pa + pitch;
        t.u00 = __ldg(pa); t.u01 = __ldg(pa + 1); t.u10 = __ldg(pb); t.u11 = __ldg(pb + 1);
    }
}

// q = R0 at the pixel; u00, u01 = top taps, u10, u11 = bottom taps of the clamped footprint
template <bool BORDER>
__device__ __forceinline__ void update_finish_core(const uint4& tq, const uint4& t00, const uint4& t01, const uint4& t10, const uint4& t11,
                                                   float tfx, float tgx, float tfy, float tsa, float tdx, float tdy, int w, int h, int x,
                                                   int y, MOut<true>& o) {
    const float gy = 1.f - tfy;
    const float a00 = tgx * gy, a01 = tfx * gy, a10 = tgx * tfy, a11 = tfx * tfy;
    // linear terms: fp32 taps, fp32 weights
    const float rb0 = a00 * __uint_as_float(t00.x) + a01 * __uint_as_float(t01.x) + a10 * __uint_as_float(t10.x) + a11 * __uint_as_float(t11.x);
    const float rb1 = a00 * __uint_as_float(t00.y) + a01 * __uint_as_float(t01.y) + a10 * __uint_as_float(t10.y) + a11 * __uint_as_float(t11.y);
    // quadratic terms: fp16 taps x fp16 weights accumulated in fp32 (the weights lose 2^-12 relative, far below the storage
    // rounding of the taps themselves)
    unsigned short w00, w01, w10, w11;
    split_h2(pack_h2_sat(a00, a01), w00, w01);
    split_h2(pack_h2_sat(a10, a11), w10, w11);
    unsigned short yy00, xx00, yy01, xx01, yy10, xx10, yy11, xx11, xy00, xy01, xy10, xy11, z;
    split_h2(t00.z, yy00, xx00); split_h2(t01.z, yy01, xx01); split_h2(t10.z, yy10, xx10); split_h2(t11.z, yy11, xx11);
    split_h2(t00.w, xy00, z); split_h2(t01.w, xy01, z); split_h2(t10.w, xy10, z); split_h2(t11.w, xy11, z);
    const float ayy = fh_fma(yy11, w11, fh_fma(yy10, w10, fh_fma(yy01, w01, fh_fma(yy00, w00, 0.f))));
    const float axx = fh_fma(xx11, w11, fh_fma(xx10, w10, fh_fma(xx01, w01, fh_fma(xx00, w00, 0.f))));
    const float axy = fh_fma(xy11, w11, fh_fma(xy10, w10, fh_fma(xy01, w01, fh_fma(xy00, w00, 0.f))));
    unsigned short qyy, qxx, qxy;
    split_h2(tq.z, qyy, qxx); split_h2(tq.w, qxy, z);
    float r4 = fh_add(qyy, ayy) * tsa;
    float r5 = fh_add(qxx, axx) * tsa;
    float r6 = fh_add(qxy, axy) * (tsa * 0.5f);
    float b2 = (__uint_as_float(tq.x) - rb0) * 0.5f;
    float b3 = (__uint_as_float(tq.y) - rb1) * 0.5f;
    if (BORDER && ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10))) {
        const float sc = border_w(x, w) * border_w(y, h);
        b2 *= sc; b3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    const float g11 = r4 * r4 + r6 * r6, g12 = (r4 + r5) * r6, g22 = r5 * r5 + r6 * r6;
    o.g01 = pack_h2_sat(g11, g12);
    o.g2 = f2h_sat(g22);
    unsigned short hg11, hg12;
    split_h2(o.g01, hg11, hg12);
    const float q11 = h2f(hg11), q12 = h2f(hg12), q22 = h2f(o.g2);
    // h = A (b0 - b1)/2 + G d with the ROUNDED G (see the note on consistent rounding above)
    o.h1 = fmaf(q12, tdx, fmaf(q11, tdy, fmaf(r6, b3, r4 * b2)));
    o.h2 = fmaf(q22, tdx, fmaf(q12, tdy, fmaf(r5, b3, r6 * b2)));
}

template <bool BORDER>
__device__ __forceinline__ void update_finish_h(const UpdTaps& t, int w, int h, int x, int y, MOut<true>& o) {
    update_finish_core<BORDER>(t.q, t.u00, t.u01, t.u10, t.u11, t.fx, t.gx, t.fy, t.sa, t.dx, t.dy, w, h, x, y, o);
}

// one-piece form (runtime-parameter kernel, tile tails that also write flow)
template <bool BORDER = true>
__device__ __forceinline__ void update_px_h(const uint4* __restrict__ R0, const uint4* __restrict__ R1, unsigned pitch,
                                            int w, int h, int x, int y, float dx, float dy, MOut<true>& o) {
    UpdTaps t;
    update_issue_h<true>(R0 + (unsigned)y * pitch + (unsigned)x, R1, pitch, w, h, x, y, dx, dy, t);
    update_finish_h<BORDER>(t, w, h, x, y, o);
}

// R0/R1 of pair p in the frame ring (p < nslots, slot0 < nslots: one conditional subtract instead of an integer modulo).
__device__ __forceinline__ int ring_slot(int slot0, int p, int nslots) {
    const int s = slot0 + p;
    return s >= nslots ? s - nslots : s;
}

// Layout-agnostic front end: RH = packed pixels (slot_stride counts uint4 pixels), else fp32 planes (slot_stride in floats).
// BORDER = false: the caller guarantees the pixel lies outside the 5-px attenuation ring (interior tiles).
template <bool RH, bool BORDER = true>
__device__ __forceinline__ void update_px_any(const void* R0, const void* R1, unsigned plane, unsigned pitch, int w, int h,
                                              int x, int y, float dx, float dy, MOut<RH>& o);
template <> __device__ __forceinline__ void update_px_any<true, true>(const void* R0, const void* R1, unsigned, unsigned pitch, int w, int h, int x, int y, float dx, float dy, MOut<true>& o) {
    update_px_h<true>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), pitch, w, h, x, y, dx, dy, o);
}
template <> __device__ __forceinline__ void update_px_any<true, false>(const void* R0, const void* R1, unsigned, unsigned pitch, int w, int h, int x, int y, float dx, float dy, MOut<true>& o) {
    update_px_h<false>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), pitch, w, h, x, y, dx, dy, o);
}
template <> __device__ __forceinline__ void update_px_any<false, true>(const void* R0, const void* R1, unsigned plane, unsigned pitch, int w, int h, int x, int y, float dx, float dy, MOut<false>& o) {
    update_px<true>(static_cast<const float*>(R0), static_cast<const float*>(R1), plane, pitch, w, h, x, y, dx, dy, o.m);
}
template <> __device__ __forceinline__ void update_px_any<false, false>(const void* R0, const void* R1, unsigned plane, unsigned pitch, int w, int h, int x, int y, float dx, float dy, MOut<false>& o) {
    update_px<false>(static_cast<const float*>(R0), static_cast<const float*>(R1), plane, pitch, w, h, x, y, dx, dy, o.m);
}

template <bool RH>
__device__ __forceinline__ const void* r_slot_ptr(const void* R, size_t slot_stride, int slot) {
    if (RH) return static_cast<const uint4*>(R) + (size_t)slot * slot_stride;
    return static_cast<const float*>(R) + (size_t)slot * slot_stride;
}

// ---------------------------------------------------------------------------------------------------
// Arguments and helpers of the fused blur + solve (+ update) (+ ROI) kernels.
// ---------------------------------------------------------------------------------------------------
constexpr int kBsTW = 32, kBsTH = 8;
// Per-CTA ROI partials: sum vx, sum vy, sum mag, then the three COUNTS of non-NaN samples (np.nanmean skips NaN per
// array, optical_flow.py:185-187), padded to two float4.
constexpr int kRoiVals = 8;

struct RoiAcc {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
    __device__ __forceinline__ void add(float vx, float vy) {
        const float mg = sqrtf(vx * vx + vy * vy);
        if (vx == vx) { s0 += vx; n0 += 1.f; }
        if (vy == vy) { s1 += vy; n1 += 1.f; }
        if (mg == mg) { s2 += mg; n2 += 1.f; }
    }
};

// CTA order of the batched kernels.  A launch covers `np` pairs x ntile tiles with a 1-D grid; consecutive CTA ids run
// through the pairs of a GROUP first, then the tiles of a row, then the tile rows, then the groups.  Frame p+1's
// coefficients are R1 of pair p and R0 of pair p+1: with the pairs of a group side by side on the SMs each R tile is
// fetched from HBM once and hit in L2 by its second reader, while vertically adjacent tiles (which share the halo rows of
// M) stay only group x tiles-per-row CTAs apart.  group = 1 is the plain order (tile fastest, pair slowest).
struct TilePos { int bx, by, p; };
__device__ __forceinline__ TilePos decode_cta(unsigned id, int nbx, int nby, int np, int group) {
    const unsigned ntile = (unsigned)(nbx * nby), g = (unsigned)group;
    const unsigned full = ((unsigned)np / g) * g;                 // pairs that sit in complete groups
    unsigned gsz = g, base = 0, rel = id;
    if (id >= full * ntile) { rel = id - full * ntile; gsz = (unsigned)np - full; base = full; }
    else { base = (id / (g * ntile)) * g; rel = id % (g * ntile); }
    const unsigned t = rel / gsz;
    TilePos r;
    r.p = (int)(base + rel % gsz);
    r.bx = (int)(t % (unsigned)nbx);
    r.by = (int)(t / (unsigned)nbx);
    return r;
}

struct BlurSolveArgs {
    int np, pair_group;                                                     // batch size and CTA order (decode_cta)
    const void* M; size_t m_stride, plane_stride; int pitch, w, h;         // matrices (MView); m_stride = BYTES per pair
    // outputs (each optional)
    float2* flow; int flow_pitch; size_t flow_stride;
    void* Mout;
    const void* R; size_t slot_stride; int slot0, nslots;   // for Mout (fp32 planes, or packed fp16 pixels)
    // ROI reduction (optional): masks [n_roi][h][w] u8; axes per pair; partial [pair][roi][ncta][4]
    const uint8_t* masks; int n_roi; size_t mask_stride; int mask_pitch;
    const float* axes;  // [pair][4] = ex0, ex1, ey0, ey1
    float* partial;
    // optional: class of every (roi, tile) of this launch's tile grid -- 0 = no ROI pixel in the tile, 1 = all of the tile's
    // pixels, 2 = mixed (k_roi_tile_class); NULL = treat every tile as mixed
    const uint8_t* roi_class;
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

// CTA-wide reduction of per-thread ROI accumulators into partial[pair][roi][cta][kRoiVals] (fixed order: deterministic).
// s_red: [nwarp][8] floats.  Every thread of the CTA must call this.
__device__ __forceinline__ void roi_cta_store(const RoiAcc& acc, float* __restrict__ s_red, float* __restrict__ dst) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const float v0 = warp_sum(acc.s0), v1 = warp_sum(acc.s1), v2 = warp_sum(acc.s2);
    const float c0 = warp_sum(acc.n0), c1 = warp_sum(acc.n1), c2 = warp_sum(acc.n2);
    __syncthreads();
    if (lane == 0) {
        float* q = s_red + warp * 8;
        q[0] = v0; q[1] = v1; q[2] = v2; q[3] = c0; q[4] = c1; q[5] = c2;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float t = 0.f;
        for (int i = 0; i < nwarp; ++i) t += s_red[i * 8 + threadIdx.x];
        dst[threadIdx.x] = t;
    }
}

// block-wide ROI partial sums for one pixel per thread
__device__ __forceinline__ void roi_reduce_store(const BlurSolveArgs& a, int p, int x, int y, bool valid,
                                                 float2 fl, float* s_red /*[8][8]*/) {
    const float* ax = a.axes + p * 4;
    const float vx = fl.x * ax[0] + fl.y * ax[1];
    const float vy = fl.x * ax[2] + fl.y * ax[3];
    const int ncta = gridDim.x * gridDim.y;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;                   // callers of this helper launch 3-D grids
    for (int r = 0; r < a.n_roi; ++r) {
        RoiAcc acc;
        if (valid && a.masks[(size_t)r * a.mask_stride + (size_t)y * a.mask_pitch + x] != 0) acc.add(vx, vy);
        roi_cta_store(acc, s_red, a.partial + (((size_t)p * a.n_roi + r) * ncta + cta) * kRoiVals);
    }
}


}  // namespace bf
