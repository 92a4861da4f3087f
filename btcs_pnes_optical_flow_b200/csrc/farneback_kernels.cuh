// Farneback dense optical flow: generic (runtime-parameter) CUDA kernels for sm_100a.
// Layouts and the shared device code: farneback_common.cuh.
#pragma once
#include "farneback_common.cuh"

namespace bf {

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

struct ResizeTab {
    const int* ix; const float* ax;  // [w]
    const int* iy; const float* ay;  // [h]
};

// K3a: M = UpdateMatrices(R0, R1, flow_init).  flow_mode: 0 = zero, 1 = flow buffer, 2 = upsample coarse.
struct UpdateArgs {
    int np, pair_group;                                                    // batch size and CTA order (decode_cta)
    const void* R; size_t plane_stride, slot_stride; int slot0, nslots;   // R0 = ring slot (slot0+p), R1 = the next one
    int pitch, w, h;
    int flow_mode;
    const float2* flow; int flow_pitch; size_t flow_stride;              // mode 1: [pair][h][flow_pitch]; mode 2: coarse
    int ws, hs; float mult; ResizeTab tab;
    void* M; size_t m_stride;                                              // matrices, m_stride = BYTES per pair (MView)
    float2* flow_out; int flow_out_pitch; size_t flow_out_stride;          // optional: write flow_init
};

// Each thread handles kUpdRows rows of one column: the x-side resize table entry, the ring-slot pointers and the kernel
// parameters are fetched once per thread instead of once per pixel (the per-pixel version was ~330 SASS instructions, a
// third of them index arithmetic and constant loads; ncu, profiles/).
constexpr int kUpdRows = 4;   // 8 rows: 74 registers, slower (18.1 vs 17.3 ms per 128 pairs)

// Compact plans: the kernel uses no shared memory, so L1 is large (256 KB per SM) -- every thread first requests the lines
// its rows will gather (R0 pixel, the two tap rows of R1) into L1 and then walks its rows one after the other, L1-hit
// loads, no register double-buffering of the taps: 40 registers, 6 CTAs per SM.  (The software-pipelined form it replaces
// -- next row's gather in flight in registers, 64 registers, 4 CTAs/SM -- took 1.45 ms per 64-pair 1080p launch, this one
// 1.23 ms; the kernel is latency-bound and responds to resident warps; profiles/r2_experiments.txt.)
template <bool RH>
__global__ void __launch_bounds__(256, RH ? 6 : 5) k_update(const UpdateArgs a) {
    // block = 64 columns x (4 x kUpdRows) rows; 1-D grid in decode_cta order
    const TilePos tp = decode_cta(blockIdx.x, (a.w + 63) / 64, (a.h + 4 * kUpdRows - 1) / (4 * kUpdRows), a.np, a.pair_group);
    const int x = tp.bx * 64 + threadIdx.x;
    const int yb = (tp.by * 4 + threadIdx.y) * kUpdRows;
    const int p = tp.p;
    if (x >= a.w || yb >= a.h) return;
    const int w = a.w, h = a.h;
    const unsigned pitch = (unsigned)a.pitch, plane = (unsigned)a.plane_stride;
    const void* R0 = nullptr;
    const void* R1 = nullptr;
    const bool want_m = a.M != nullptr;
    const MView<RH> Mo(a.M, a.m_stride, p, plane, pitch, w);
    if (want_m) {
        R0 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p, a.nslots));
        R1 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p + 1, a.nslots));
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
    if (!want_m) return;
    if constexpr (RH) {
        const uint4* R0h = static_cast<const uint4*>(R0);
        const uint4* R1h = static_cast<const uint4*>(R1);
        // the thread's kUpdRows rows lie in one block of M (yb is a multiple of kUpdRows, which divides the block height):
        // rows at +256 / +512 bytes from one pair of pointers
        static_assert(kMbH % kUpdRows == 0, "a thread's rows must not straddle blocks");
        char* blk = Mo.block(x >> 7, yb >> 4);
        char* pg = blk + MView<true>::g_off(yb & 15, x & 127);
        char* ph = blk + MView<true>::h_off(yb & 15, x & 127);
        // (the always-true runtime test keeps the compiler from carrying this loop's tap coordinates into the next one:
        // merged, the two loops spill 190 bytes at 40 registers)
        if (a.np > 0) {
#pragma unroll
            for (int k = 0; k < kUpdRows; ++k) {
                const int yy = min(yb + k, h - 1);
                const int x1 = __float2int_rd((float)x + fl[k].x), y1 = __float2int_rd((float)yy + fl[k].y);
                const int cx = max(min(x1, w - 2), 0), cy = max(min(y1, h - 2), 0);
                const uint4* pa = R1h + (unsigned)cy * pitch + (unsigned)cx;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pa));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pa + pitch));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(R0h + (unsigned)yy * pitch + (unsigned)x));
            }
        }
#pragma unroll
        for (int k = 0; k < kUpdRows; ++k) {
            const int yy = min(yb + k, h - 1);
            UpdTaps T;
            update_issue_h(R0h + (unsigned)yy * pitch + (unsigned)x, R1h, pitch, w, h, x, yy, fl[k].x, fl[k].y, T);
            MOut<true> m;
            if (inner) update_finish_h<false>(T, w, h, x, yb + k, m); else update_finish_h<true>(T, w, h, x, yb + k, m);
            if (yb + k < h) MView<true>::store_at(pg + k * (kMbW * 2), ph + k * (kMbW * 4), m);
        }
    } else {
        // exact plans: the same L1 prefetch over the five fp32 planes (two tap rows each + the R0 pixel)
        if (a.np > 0) {
            const float* R0f = static_cast<const float*>(R0);
            const float* R1f = static_cast<const float*>(R1);
#pragma unroll
            for (int k = 0; k < kUpdRows; ++k) {
                const int yy = min(yb + k, h - 1);
                const int x1 = __float2int_rd((float)x + fl[k].x), y1 = __float2int_rd((float)yy + fl[k].y);
                const int cx = max(min(x1, w - 2), 0), cy = max(min(y1, h - 2), 0);
                const float* pa = R1f + (unsigned)cy * pitch + (unsigned)cx;
                const float* pq = R0f + (unsigned)yy * pitch + (unsigned)x;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pa + (size_t)c * plane));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pa + (size_t)c * plane + pitch));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pq + (size_t)c * plane));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kUpdRows; ++k) {
            const int y = yb + k;
            if (y >= h) break;
            MOut<false> m;
            if (inner) update_px_any<false, false>(R0, R1, plane, pitch, w, h, x, y, fl[k].x, fl[k].y, m);
            else update_px_any<false, true>(R0, R1, plane, pitch, w, h, x, y, fl[k].x, fl[k].y, m);
            Mo.store(y, x, m);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K3b (generic): flow = Solve(Blur(M)) [+ M' = UpdateMatrices(flow)] [+ projection and ROI sums].
// Tile 32 x 8 outputs, 256 threads.  Phase 1: vertical window sums global -> shared (one thread per
// (channel, column), exact first window then short-history sliding: no long-range cancellation).
// Phase 2: horizontal sums from shared, accurate 2x2 solve.  Phase 3: fused tail.
// ---------------------------------------------------------------------------------------------------
template <bool RH>
__global__ void __launch_bounds__(256) k_blur_solve_generic(const BlurSolveArgs a, const WinCoef wc) {
    __shared__ float V[5][kBsTH][kBsTW + 2 * kMaxWinHalf];
    __shared__ float s_red[8 * 8];
    const int m = wc.m;
    const int x0 = blockIdx.x * kBsTW, y0 = blockIdx.y * kBsTH;
    const int p = blockIdx.z;
    const MView<RH> Mp(const_cast<void*>(a.M), a.m_stride, p, (unsigned)a.plane_stride, (unsigned)a.pitch, a.w);
    const int tw = kBsTW + 2 * m;
    const int w = a.w, h = a.h;
    for (int it = threadIdx.x; it < 5 * tw; it += blockDim.x) {
        const int c = it / tw, tx = it - c * tw;
        const int gx = min(max(x0 - m + tx, 0), w - 1);
        auto at = [&](int row) { return Mp.load(c, row, gx); };
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
            MOut<RH> mm;
            update_px_any<RH, true>(R0, R1, (unsigned)a.plane_stride, (unsigned)a.pitch, w, h, x, y, fl.x, fl.y, mm);
            MView<RH>(a.Mout, a.m_stride, p, (unsigned)a.plane_stride, (unsigned)a.pitch, w).store(y, x, mm);
        }
    }
    if (a.partial) roi_reduce_store(a, p, x, y, valid, fl, s_red);
}

// Projection + ROI partial sums of an existing flow buffer (iterations == 0 path; same partial layout).
__global__ void __launch_bounds__(256) k_roi_from_flow(const BlurSolveArgs a) {
    __shared__ float s_red[8 * 8];
    const int tx = threadIdx.x & (kBsTW - 1), ty = threadIdx.x / kBsTW;
    const int x = blockIdx.x * kBsTW + tx, y = blockIdx.y * kBsTH + ty;
    const int p = blockIdx.z;
    const bool valid = (x < a.w) && (y < a.h);
    float2 fl = make_float2(0.f, 0.f);
    if (valid) fl = a.flow[(size_t)p * a.flow_stride + (size_t)y * a.flow_pitch + x];
    roi_reduce_store(a, p, x, y, valid, fl, s_red);
}

// Finalise the ROI means: out[roi][t_first+p][3] = sums / counts (np.nanmean of each array: NaN samples were skipped, a
// ROI without samples gives NaN).  One warp per (pair, roi); lanes stride over the per-CTA partials and accumulate in
// double, then a fixed-order shuffle tree: deterministic run to run.
__global__ void k_roi_finalize(const float* __restrict__ partial, int n_pairs, int n_roi, int ncta,
                               const double* __restrict__ ex, const double* __restrict__ ey, int t_first,
                               float* __restrict__ out, int T) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n_pairs * n_roi) return;
    const int p = i / n_roi, r = i - p * n_roi;
    const float4* q = reinterpret_cast<const float4*>(partial + (size_t)i * ncta * kRoiVals);
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int c = lane; c < ncta; c += 32) {
        const float4 u = q[2 * c], v = q[2 * c + 1];
        s[0] += u.x; s[1] += u.y; s[2] += u.z; s[3] += u.w; s[4] += v.x; s[5] += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 6; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    }
    if (lane != 0) return;
    const int t = t_first + p;
    float* o = out + ((size_t)r * T + t) * 3;
    const bool ok = isfinite(ex[2 * t]) && isfinite(ex[2 * t + 1]) && isfinite(ey[2 * t]) && isfinite(ey[2 * t + 1]);
    const float nanv = __int_as_float(0x7fc00000);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = (ok && s[3 + k] > 0.0) ? (float)(s[k] / s[3 + k]) : nanv;
}

// Class of every (roi, tile) for tiles of tw x th pixels: 0 = the mask is zero on the whole tile, 1 = non-zero on all of its
// in-image pixels, 2 = mixed.  Once per series call; lets the last iteration skip the per-pixel mask loads (a full-frame ROI,
// config C2, is class 1 everywhere; a small ROI leaves most tiles at class 0).  One warp per tile.
__global__ void __launch_bounds__(128) k_roi_tile_class(const uint8_t* __restrict__ masks, int n_roi, int W, int H, int tw, int th,
                                                        int nbx, int nby, uint8_t* __restrict__ cls) {
    const int tile = blockIdx.x * 4 + (threadIdx.x >> 5), roi = blockIdx.y, lane = threadIdx.x & 31;
    if (tile >= nbx * nby) return;
    const int bx = tile % nbx, by = tile / nbx;
    const int x0 = bx * tw, y0 = by * th, x1 = min(x0 + tw, W), y1 = min(y0 + th, H);
    const uint8_t* m = masks + (size_t)roi * W * H;
    bool any = false, all = true;
    for (int y = y0; y < y1; ++y)
        for (int x = x0 + lane; x < x1; x += 32) {
            const bool on = m[(size_t)y * W + x] != 0;
            any |= on; all &= on;
        }
    any = __any_sync(0xffffffffu, any);
    all = __all_sync(0xffffffffu, all);
    if (lane == 0) cls[(size_t)roi * nbx * nby + tile] = all ? 1 : (any ? 2 : 0);
}

// OPTFLOW_USE_INITIAL_FLOW: the caller's full-resolution flow resized to the coarsest scale with cv2's INTER_AREA weights
// (tables built on the host: for destination index d the source taps ofs[d] .. ofs[d+1]-1) and multiplied by the scale.
struct AreaTab { const int* ofs; const int* idx; const float* wgt; };
__global__ void __launch_bounds__(256) k_resize_area_flow(const float2* __restrict__ src, int W, size_t src_stride, AreaTab tx, AreaTab ty,
                                                          int w, int h, float mult, float2* __restrict__ dst, int dst_pitch, size_t dst_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, p = blockIdx.z;
    if (x >= w || y >= h) return;
    const float2* s = src + (size_t)p * src_stride;
    float ax = 0.f, ay = 0.f;
    for (int j = ty.ofs[y]; j < ty.ofs[y + 1]; ++j) {
        const float2* row = s + (size_t)ty.idx[j] * W;
        float rx = 0.f, ry = 0.f;
        for (int i = tx.ofs[x]; i < tx.ofs[x + 1]; ++i) {
            const float2 v = row[tx.idx[i]];
            const float a = tx.wgt[i];
            rx += v.x * a; ry += v.y * a;
        }
        ax += rx * ty.wgt[j]; ay += ry * ty.wgt[j];
    }
    dst[(size_t)p * dst_stride + (size_t)y * dst_pitch + x] = make_float2(ax * mult, ay * mult);
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
