// Specialised (compile-time poly_n) polynomial expansion, pyramid and colour-conversion kernels.
#pragma once
#include <cstdlib>
#include <cstring>

#include "farneback_kernels.cuh"
#include "farneback_tile.cuh"

namespace bf {

// ---------------------------------------------------------------------------------------------------
// k_polyexp<N>: separable polynomial expansion with compile-time poly_n.  Tile 128 x 24, 256 threads.
// 72 % of the first version's 133 instructions per pixel were FFMA / FADD / FMUL and the rest per-element overhead; both
// passes now work on packed fp32 pairs (FFMA2 / FADD2, farneback_common.cuh -- same issue cost as two scalar operations, but
// one load, address and loop step per pair):
//   vertical pass   task = (column pair, 8-row group): 64-bit loads straight from global (rows clamped = replicate), the
//                   8 + 2N row pairs live in registers; per tap one packed add, one packed subtract, three packed FMAs
//   shared memory   plane P01 = (t0, t1) interleaved per column, stored as 16-byte chunks of two columns, even chunks
//                   first and odd chunks from chunk HOFF on (HOFF = 4 mod 8: conflict-free STS.128 here and LDS.128 with a
//                   lane stride of two chunks in the horizontal pass); plane P2 = t2 plain
//   horizontal pass thread = 4 adjacent outputs, scatter form: every loaded column is multiplied into the outputs it reaches
//                   -- (b1, b3) += g (t0, t1), (b2, b6) += +-xg (t0, t1) as pairs, b4 += xxg t0 and b5 += g t2 scalar -- so
//                   no window is held in registers; the 4 packed pixels leave as two 256-bit stores
// Sums run in row / column order (cv2: centre, then symmetric pairs): an fp32 reordering of the coefficients, ~1e-7 rel.
// ---------------------------------------------------------------------------------------------------
constexpr int kPeFastTH = 24;
template <int N>
struct FastPeCfg {
    static constexpr int HALO = (N + 1) & ~1;                   // even: column pairs start on even columns
    static constexpr int D = HALO - N;
    static constexpr int NCOL = kFbTW + 2 * HALO;
    static constexpr int NPAIR = NCOL / 2;                      // column pairs = 16-byte chunks of P01 per row
    static constexpr int NE = (NPAIR + 1) / 2;                  // even chunks
    static constexpr int HOFF = NE + ((12 - NE % 8) % 8);       // first odd chunk: >= NE and = 4 (mod 8)
    static constexpr int ROWCH = HOFF + NPAIR / 2;              // chunks per P01 row
    static constexpr int VP = (NCOL + 3) / 4 * 4;               // P2 row pitch (floats)
    static constexpr int RPG = 8, RG = kPeFastTH / RPG;             // rows per task, row groups
    static constexpr int NCH = (D + 2 * N + 3) / 4 + 1;         // 4-column groups read by a thread of the horizontal pass
    static constexpr int P01_FLOATS = kPeFastTH * ROWCH * 4;
    static constexpr size_t SMEM = (size_t)(P01_FLOATS + kPeFastTH * VP) * sizeof(float);
    static_assert(NCOL % 4 == 0 && HOFF % 8 == 4 && HOFF >= NE, "chunk layout");
    static_assert(4 * 31 + 4 * NCH <= NCOL, "horizontal pass must stay inside the tile row");
    static_assert(RG * NPAIR <= 256, "one vertical task per thread");
};

// two adjacent 16-byte pixels as one 256-bit store; p must be 32-byte aligned
__device__ __forceinline__ void store_px2(uint4* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
                 "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

template <int N, bool RH>
__global__ void __launch_bounds__(256, N <= 5 ? 4 : 3) k_polyexp(const float* __restrict__ I, int pitch, size_t frame_stride, int w,
                                                                 int h, void* __restrict__ Rv, size_t plane_stride,
                                                                 size_t slot_stride, int slot0, int nslots, const PolyCoef pc) {
    using C = FastPeCfg<N>;
    extern __shared__ __align__(16) float smem[];
    float* const P01 = smem;
    float* const P2 = smem + C::P01_FLOATS;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFbTW, y0 = blockIdx.y * kPeFastTH, f = blockIdx.z;
    const float* img = I + (size_t)f * frame_stride;

    // ---------------- vertical pass: column pairs ----------------
    if (tid < C::RG * C::NPAIR) {
        const int rg = tid / C::NPAIR, cp = tid - rg * C::NPAIR;
        // tile columns 2cp, 2cp + 1 = image columns x, x + 1 (x even); columns outside the image replicate the border
        // column: the pair is loaded from the even column xe and the halves are picked after the sums (the filter is linear)
        const int x = x0 - C::HALO + 2 * cp;
        const int cl = min(max(x, 0), w - 1), ch = min(max(x + 1, 0), w - 1);
        const int xe = cl & ~1;
        const bool lo_y = (cl & 1) != 0, hi_y = (ch - xe) != 0;
        const bool straight = !lo_y && hi_y;                        // both columns inside: no selection
        const int yb = y0 + rg * C::RPG - N;
        f32x2_t v[C::RPG + 2 * N];
#pragma unroll
        for (int i = 0; i < C::RPG + 2 * N; ++i) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(img + (size_t)min(max(yb + i, 0), h - 1) * pitch + xe));
            v[i] = pk2(t.x, t.y);
        }
        f32x2_t g2[N + 1], xg2[N + 1], xxg2[N + 1];
#pragma unroll
        for (int k = 0; k <= N; ++k) { g2[k] = pk2(pc.g[k], pc.g[k]); xg2[k] = pk2(pc.xg[k], pc.xg[k]); xxg2[k] = pk2(pc.xxg[k], pc.xxg[k]); }
        const f32x2_t m1 = pk2(-1.f, -1.f);
        const int chunk = cp;
        float* d01 = P01 + ((size_t)(rg * C::RPG) * C::ROWCH + (chunk >> 1) + (chunk & 1) * C::HOFF) * 4;
        float* d2 = P2 + (size_t)(rg * C::RPG) * C::VP + 2 * cp;
#pragma unroll
        for (int j = 0; j < C::RPG; ++j) {
            f32x2_t t0 = mul2(v[j + N], g2[0]), t1, t2;
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                const f32x2_t up = v[j + N - k], dn = v[j + N + k];
                const f32x2_t pp = add2(up, dn);
                const f32x2_t dd = fma2(up, m1, dn);                // dn - up (exact: the product by -1 does not round)
                t0 = fma2(pp, g2[k], t0);
                t1 = k == 1 ? mul2(dd, xg2[k]) : fma2(dd, xg2[k], t1);
                t2 = k == 1 ? mul2(pp, xxg2[k]) : fma2(pp, xxg2[k], t2);
            }
            float2 a = up2(t0), b = up2(t1), c = up2(t2);
            if (!straight) {
                a = make_float2(lo_y ? a.y : a.x, hi_y ? a.y : a.x);
                b = make_float2(lo_y ? b.y : b.x, hi_y ? b.y : b.x);
                c = make_float2(lo_y ? c.y : c.x, hi_y ? c.y : c.x);
            }
            *reinterpret_cast<float4*>(d01 + (size_t)j * (C::ROWCH * 4)) = make_float4(a.x, b.x, a.y, b.y);
            *reinterpret_cast<float2*>(d2 + (size_t)j * C::VP) = c;
        }
    }
    __syncthreads();

    // ---------------- horizontal pass + coefficient store ----------------
    const int g = tid & 31, rb = tid >> 5;
    const int slot = (slot0 + f) % nslots;
    float* Rb = RH ? nullptr : static_cast<float*>(Rv) + (size_t)slot * slot_stride;
    uint4* Rh = RH ? static_cast<uint4*>(Rv) + (size_t)slot * slot_stride : nullptr;
    const bool wide = (reinterpret_cast<uintptr_t>(Rh) & 31) == 0;      // pitch and x are multiples of 4 pixels: rows stay 32-byte aligned
    f32x2_t g2[N + 1], pxg2[N + 1], nxg2[N + 1];
#pragma unroll
    for (int k = 0; k <= N; ++k) { g2[k] = pk2(pc.g[k], pc.g[k]); pxg2[k] = pk2(pc.xg[k], pc.xg[k]); nxg2[k] = pk2(-pc.xg[k], -pc.xg[k]); }
#pragma unroll 1
    for (int k = 0; k < kPeFastTH / 8; ++k) {
        const int r = rb + 8 * k;
        const int y = y0 + r, x = x0 + 4 * g;
        // thread g: tile columns 4g .. 4g + 4 NCH - 1 = chunks 2g + i; output j sits at tile column HALO + 4g + j
        const ulonglong2* row01 = reinterpret_cast<const ulonglong2*>(P01 + (size_t)r * (C::ROWCH * 4)) + g;
        const float4* row2 = reinterpret_cast<const float4*>(P2 + (size_t)r * C::VP) + g;
        f32x2_t b13[4], b26[4];
        float b4[4], b5[4];
#pragma unroll
        for (int i = 0; i < 2 * C::NCH; ++i) {
            const ulonglong2 q = row01[(i >> 1) + (i & 1) * C::HOFF];
            float4 q2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((i & 1) == 0) q2 = row2[i >> 1];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int l = 2 * i + hf;                             // tile column 4g + l
                const f32x2_t v01 = hf ? q.y : q.x;
                const float v0 = up2(v01).x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = l - (C::HALO + j);                  // column offset from output j
                    if (d < -N || d > N) continue;
                    const int ad = d < 0 ? -d : d;
                    if (d == -N) {
                        b13[j] = mul2(v01, g2[N]);
                        b26[j] = mul2(v01, nxg2[N]);
                        b4[j] = v0 * pc.xxg[N];
                    } else {
                        b13[j] = fma2(v01, g2[ad], b13[j]);
                        if (d != 0) {
                            b26[j] = fma2(v01, d < 0 ? nxg2[ad] : pxg2[ad], b26[j]);
                            b4[j] = fmaf(v0, pc.xxg[ad], b4[j]);
                        }
                    }
                }
            }
            // t2 of the same 4 columns (every second chunk step)
            if ((i & 1) == 0) {
                const float vv[4] = {q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int d = 4 * (i >> 1) + e - (C::HALO + j);
                        if (d < -N || d > N) continue;
                        const int ad = d < 0 ? -d : d;
                        b5[j] = d == -N ? vv[e] * pc.g[N] : fmaf(vv[e], pc.g[ad], b5[j]);
                    }
                }
            }
        }
        float o[5][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 s13 = up2(b13[j]), s26 = up2(b26[j]);
            const float b1 = s13.x, b3 = s13.y, b2 = s26.x, b6 = s26.y;
            o[0][j] = b3 * pc.ig11;
            o[1][j] = b2 * pc.ig11;
            o[2][j] = fmaf(b1, pc.ig03, b5[j] * pc.ig33);
            o[3][j] = fmaf(b1, pc.ig03, b4[j] * pc.ig33);
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
        // the level's descriptor and the four output row pointers are taken once per level (indexed reads of the kernel
        // parameter and 64-bit address arithmetic per output column were half of this kernel's instructions; ncu)
        const PyrLevelDesc d = pa.lv[l];
        const int rad = d.ksize >> 1, ksize = d.ksize, lw = d.w;
        const float* __restrict__ kern = d.kern;
        const int* __restrict__ lix = d.ix;
        const float* __restrict__ lax = d.ax;
        float* trow[kPyrRows];
#pragma unroll
        for (int rr = 0; rr < kPyrRows; ++rr)                      // rows past the frame write (again) to the last valid row
            trow[rr] = d.tmp + (size_t)f * d.tmp_frame_stride + (size_t)min(r0 + rr, H - 1) * d.tmp_pitch;
        for (int x = threadIdx.x; x < lw; x += 256) {
            const int i0 = lix[x];
            const float a = lax[x];
            // the 4 rows of a source pixel are two packed fp32 pairs: one FFMA2 per pair and tap (coefficient broadcast)
            const ulonglong2* srow2 = reinterpret_cast<const ulonglong2*>(srow4);
            const f32x2_t z2 = pk2(0.f, 0.f);
            f32x2_t b0l = z2, b0h = z2, b1l = z2, b1h = z2;
            if (i0 - rad >= 0 && i0 + 1 + rad < W) {
                int i = i0 - rad;
                ulonglong2 prev = srow2[pyr_slot(i)];
                for (int j = 0; j < ksize; ++j) {
                    const float kj = __ldg(kern + j);
                    const f32x2_t k2 = pk2(kj, kj);
                    ++i;
                    const ulonglong2 nxt = srow2[pyr_slot(i)];
                    b0l = fma2(prev.x, k2, b0l); b0h = fma2(prev.y, k2, b0h);
                    b1l = fma2(nxt.x, k2, b1l); b1h = fma2(nxt.y, k2, b1h);
                    prev = nxt;
                }
            } else {
                const int i1 = min(i0 + 1, W - 1);
                for (int j = 0; j < ksize; ++j) {
                    const float kj = __ldg(kern + j);
                    const f32x2_t k2 = pk2(kj, kj);
                    const ulonglong2 v0 = srow2[pyr_slot(reflect101(i0 - rad + j, W))], v1 = srow2[pyr_slot(reflect101(i1 - rad + j, W))];
                    b0l = fma2(v0.x, k2, b0l); b0h = fma2(v0.y, k2, b0h);
                    b1l = fma2(v1.x, k2, b1l); b1h = fma2(v1.y, k2, b1h);
                }
            }
            const float2 b0a = up2(b0l), b0b = up2(b0h), b1a = up2(b1l), b1b = up2(b1h);
            const float4 b0 = make_float4(b0a.x, b0a.y, b0b.x, b0b.y), b1 = make_float4(b1a.x, b1a.y, b1b.x, b1b.y);
            const float o[kPyrRows] = {(a != 0.f) ? (b0.x * (1.f - a) + b1.x * a) : b0.x, (a != 0.f) ? (b0.y * (1.f - a) + b1.y * a) : b0.y,
                                       (a != 0.f) ? (b0.z * (1.f - a) + b1.z * a) : b0.z, (a != 0.f) ? (b0.w * (1.f - a) + b1.w * a) : b0.w};
#pragma unroll
            for (int rr = 0; rr < kPyrRows; ++rr) trow[rr][x] = o[rr];
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

// R: 128-bit stores; I: the vertical pass loads column pairs (64-bit)
inline bool polyexp_fast_aligned(const void* R, size_t plane_stride, size_t slot_stride, const void* I, size_t frame_stride) {
    return aligned16(R) && (plane_stride % 4) == 0 && (slot_stride % 4) == 0 && (reinterpret_cast<uintptr_t>(I) & 7) == 0 && (frame_stride % 2) == 0;
}

template <int N, bool RH>
inline void launch_polyexp_fast_n(const float* I, int pitch, size_t frame_stride, int w, int h, void* R,
                                  size_t plane_stride, size_t slot_stride, int slot0, int nslots, int nf,
                                  const PolyCoef& pc, cudaStream_t st) {
    using C = FastPeCfg<N>;
    cudaFuncSetAttribute(k_polyexp<N, RH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    dim3 g((w + kFbTW - 1) / kFbTW, (h + kPeFastTH - 1) / kPeFastTH, nf);
    k_polyexp<N, RH><<<g, 256, C::SMEM, st>>>(I, pitch, frame_stride, w, h, R, plane_stride, slot_stride, slot0, nslots, pc);
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
