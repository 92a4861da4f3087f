// Specialised (compile-time window / poly_n) kernels for the configurations BASELINE.json names.
//
// k_blur_solve_box<MH>: flow = Solve(BoxBlur_{2MH+1}(M)) fused with M' = UpdateMatrices(flow) and/or the
// body-axis projection + ROI partial sums (reference call site optical_flow.py:173, reduction
// optical_flow.py:176-187).  One CTA = 128 x TH output pixels (TH = 16 by default), 256 threads, 4 CTAs per SM.
//   phase 0  three tensor-map prefetches (UTMAPF) bring the M tile and the R0 / R1 blocks of the tile into L2.
//   phase 1  vertical window sums, global -> shared.  One thread per (channel, 4-column group): 64-bit (fp16 M) or 128-bit
//            coalesced loads, the 2MH+1 row window lives in registers (fully unrolled ring; raw fp16 rows consumed by
//            FHADD on compact plans), exact first window then add-new/subtract-old with a history bounded by the TH+2MH
//            rows of the tile: no long-range cancellation, and all-zero (static) regions stay exactly zero.
//   phase 2  horizontal window sums from shared with conflict-free LDS.128 (lane stride 16 B), 4 outputs per
//            thread sharing the common partial sum (no subtraction), then the 2x2 solve with Kahan-accurate
//            determinants.  The 1/winsize^2 scale is folded into the regulariser (reg = 1e-3 * winsize^4).
//   phase 3  flow is transposed through shared memory so that lanes own consecutive pixels again: coalesced R0
//            loads, bilinear R1 gather issued one pixel ahead, M' stores; ROI sums reduced per CTA (deterministic partials).
#pragma once
#include <cstdlib>
#include <cstring>

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "farneback_kernels.cuh"

namespace bf {

constexpr int kFbTW = 128, kFbTH = 32;

template <int MH, int TH = kFbTH>
struct FastBoxCfg {
    static constexpr int HALO = (MH + 3) / 4 * 4;
    static constexpr int D = HALO - MH;                        // unused leading columns in the halo
    static constexpr int NC4 = (kFbTW + 2 * HALO) / 4;         // float4 columns per tile row
    static constexpr int VP = kFbTW + 2 * HALO + 4;            // shared row pitch (floats, multiple of 4)
    static constexpr int WIN = 2 * MH + 1;
    static constexpr int NCH = (D + 2 * MH + 3) / 4 + 1;       // float4 chunks a 4-output group reads
    static constexpr int V_FLOATS = 5 * TH * VP;
    static constexpr int RG = TH / 8;                          // row groups per thread in phases 2/3
    static constexpr int PF = TH >= 32 ? 8 : 4;                // register prefetch depth in phase 1
    static constexpr int CTAS = TH >= 32 ? 2 : (TH >= 24 ? 3 : 4);   // CTAs per SM the shared-memory footprint allows
    static constexpr size_t SMEM = (size_t)(V_FLOATS + 64) * sizeof(float);
    static_assert(MH >= 2 && MH <= 16, "half window out of range for the fast path");
    static_assert(TH % 8 == 0, "tile height must be a multiple of 8");
    static_assert((size_t)TH * kFbTW * sizeof(float2) <= (size_t)V_FLOATS * sizeof(float), "F must fit in V");
};

// Debug builds (-DBF_TRACE, tools/trace_phases.py): thread 0 of every CTA of k_blur_solve_box stamps the SM clock at the
// phase boundaries so that phase durations and the overlap of co-resident CTAs can be read off directly.
#ifdef BF_TRACE
__device__ unsigned long long* bf_trace_buf = nullptr;
__device__ __forceinline__ void trace_stamp(int slot, bool on) {
    if (threadIdx.x == 0 && on && bf_trace_buf) {
        const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        unsigned long long t;
        if (slot == 0) {
            unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            bf_trace_buf[cta * 8 + 6] = t;
            bf_trace_buf[cta * 8 + 7] = sm;
        }
        bf_trace_buf[cta * 8 + slot] = clock64();
    }
}
#define BF_TRACE_STAMP(k) trace_stamp(k, a.Mout != nullptr)   // launches with the update tail only
#else
#define BF_TRACE_STAMP(k)
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// One instruction requests a whole contiguous span into L2 through the bulk-copy unit; the per-line form above costs one
// L1 tag cycle per 128-byte line (~1800 lines per CTA of k_blur_solve_box: ~2700 cycles, 7 % of a CTA's lifetime in the
// phase trace, profiles/).  p must be 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void prefetch_l2_span(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// A whole 4-D box (tile of M with its halo: x, y, channel, pair; block of R: 4 words, x, y, ring slot) requested into L2
// by ONE instruction through a tensor map; parts of the box outside the tensor are skipped by the hardware.
__device__ __forceinline__ void prefetch_l2_box(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// Tensor maps of one launch of k_blur_solve_box (built per scale at plan creation, bf::encode_tile_maps in btcsflow.cu):
// m = input matrices {x, y, 5 channels, pair}, box (TW + 2 HALO) x (TH + 2 MH) x 5 x 1; r0 / r1 = packed R ring
// {128 words = 32 pixels, x / 32, y, slot}, boxes 128 x 4 x TH x 1 and 128 x 6 x (TH + 4) x 1.
struct TileMaps { CUtensorMap m, r0, r1; };

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// One row of a float4 column as it sits in the register window.  fp32 matrices: the float4 itself.  fp16 matrices: the
// four raw halves (half the registers), consumed by the mixed-precision add of sm_100a (FHADD: f32 + f16 -> f32 in one
// instruction, exact conversion included) so that no separate fp16 -> fp32 conversions are issued.
__device__ __forceinline__ float fh_add(unsigned short h, float s) { float d; asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(s)); return d; }
__device__ __forceinline__ float fh_sub(unsigned short h, float s) { float d; asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(s)); return d; }   // h - s
struct HalfRow {
    uint2 v;
    __device__ __forceinline__ void split(unsigned short h[4]) const {
        asm("mov.b32 {%0, %1}, %2;" : "=h"(h[0]), "=h"(h[1]) : "r"(v.x));
        asm("mov.b32 {%0, %1}, %2;" : "=h"(h[2]), "=h"(h[3]) : "r"(v.y));
    }
};
__device__ __forceinline__ float4 row_load(const float* p, float4*) { return m_load4(p); }
__device__ __forceinline__ HalfRow row_load(const __half* p, HalfRow*) { return HalfRow{__ldg(reinterpret_cast<const uint2*>(p))}; }
__device__ __forceinline__ float4 row_first(const float4& r) { return r; }
__device__ __forceinline__ float4 row_first(const HalfRow& r) {
    unsigned short h[4];
    r.split(h);
    return make_float4(fh_add(h[0], 0.f), fh_add(h[1], 0.f), fh_add(h[2], 0.f), fh_add(h[3], 0.f));
}
__device__ __forceinline__ void row_add(float4& s, const float4& r) { s = f4add(s, r); }
__device__ __forceinline__ void row_add(float4& s, const HalfRow& r) {
    unsigned short h[4];
    r.split(h);
    s.x = fh_add(h[0], s.x); s.y = fh_add(h[1], s.y); s.z = fh_add(h[2], s.z); s.w = fh_add(h[3], s.w);
}
// window slides one row: s += nv - ov
__device__ __forceinline__ void row_slide(float4& s, const float4& nv, const float4& ov) { s = f4add(s, f4sub(nv, ov)); }
__device__ __forceinline__ void row_slide(float4& s, const HalfRow& nv, const HalfRow& ov) {
    unsigned short n[4], o[4];
    nv.split(n); ov.split(o);
    s.x = fh_sub(n[0], fh_sub(o[0], s.x));                   // nv - (ov - s)
    s.y = fh_sub(n[1], fh_sub(o[1], s.y));
    s.z = fh_sub(n[2], fh_sub(o[2], s.z));
    s.w = fh_sub(n[3], fh_sub(o[3], s.w));
}
template <typename MT> struct RowOf { using type = float4; };
template <> struct RowOf<__half> { using type = HalfRow; };

// Vertical (2MH+1)-row box sums of one float4 column for TH consecutive output rows; register ring window, software
// prefetch PF rows ahead.  ROWS_IN: the tile's rows (with halo) lie inside the image, `src` already points at the first
// halo row.  Otherwise rows are clamped to [0, h-1] (replicate) starting from row `yb`.
// mode (EDGE only): 0 = group inside the image, 1 = left of it (splat lane 0 of the first group), 2 = right of it (splat
// lane kl of the last group), 3 = the last group of an image whose width is not a multiple of 4 (lanes above kl take lane kl).
template <int MH, bool ROWS_IN, bool EDGE, int TH, int PF, typename MT>
__device__ __forceinline__ void vertical_box_sums(const MT* __restrict__ src, unsigned pitch, int yb, int h, int mode, int kl,
                                                  float* __restrict__ dst, int vp) {
    constexpr int WIN = 2 * MH + 1, NROW = TH + 2 * MH;
    using Row = typename RowOf<MT>::type;
    auto ld = [&](int i) -> Row {
        if (ROWS_IN) return row_load(src + (unsigned)i * pitch, (Row*)nullptr);
        const int r = min(max(yb + i, 0), h - 1);
        return row_load(src + (unsigned)r * pitch, (Row*)nullptr);
    };
    auto st = [&](int j, const float4& s) {
        if (!EDGE) {                                      // compile-time: interior columns store the sums as they are
            *reinterpret_cast<float4*>(dst + j * vp) = s;
        } else if (mode == 0) {
            *reinterpret_cast<float4*>(dst + j * vp) = s;
        } else {
            const float e = (mode == 1 || kl == 0) ? s.x : (kl == 1 ? s.y : (kl == 2 ? s.z : s.w));
            const bool keep = mode == 3;
            *reinterpret_cast<float4*>(dst + j * vp) = make_float4(keep ? s.x : e, (keep && kl >= 1) ? s.y : e, (keep && kl >= 2) ? s.z : e, e);
        }
    };
    Row win[WIN];
#pragma unroll
    for (int i = 0; i < WIN; ++i) win[i] = ld(i);
    float4 s = row_first(win[0]);
#pragma unroll
    for (int i = 1; i < WIN; ++i) row_add(s, win[i]);
    st(0, s);
    Row pre[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) pre[i] = ld(WIN + i);
#pragma unroll
    for (int j = 1; j < TH; ++j) {
        const Row nv = pre[(j - 1) % PF];
        if (j - 1 + PF + WIN < NROW) pre[(j - 1) % PF] = ld(WIN + j - 1 + PF);
        const Row ov = win[(j - 1) % WIN];
        row_slide(s, nv, ov);
        win[(j - 1) % WIN] = nv;
        st(j, s);
    }
}

// Lines of R that a block of rows [ya, ya+NROWS) x columns [x0, x0+128) will touch in its update tail: R0 under the
// block, R1 within +-2 rows / +-32 columns (larger flows simply miss).  Requested into L2 ahead of use.  The index math
// is kept to shifts and constant divisions: these loops used to cost ~25 instructions per pixel (ncu, profiles/).
template <bool RH, int NROWS>
__device__ __forceinline__ void prefetch_r_block(const void* R0v, const void* R1v, unsigned plane, unsigned pitch, int w, int h,
                                                 int x0, int ya, int tid, int nthreads) {
    if (RH) {
        const uint4* R0 = static_cast<const uint4*>(R0v);
        const uint4* R1 = static_cast<const uint4*>(R1v);
        const int xmax = max(w - 1, 0);                                  // 8 pixels per 128-byte line, 16 lines per row
        for (int e = tid; e < NROWS * 16; e += nthreads) {
            const int yy = min(ya + (e >> 4), h - 1), xx = min(x0 + (e & 15) * 8, xmax);
            prefetch_l2(R0 + (unsigned)yy * pitch + (unsigned)xx);
        }
        for (int e = tid; e < (NROWS + 4) * 24; e += nthreads) {
            const int r = e / 24, l = e - r * 24;
            const int yy = min(max(ya - 2 + r, 0), h - 1), xx = min(max(x0 - 32 + l * 8, 0), xmax);
            prefetch_l2(R1 + (unsigned)yy * pitch + (unsigned)xx);
        }
    } else {
        const float* R0 = static_cast<const float*>(R0v);
        const float* R1 = static_cast<const float*>(R1v);
        const int xmax = (int)pitch - 32;
        for (int e = tid; e < NROWS * 4; e += nthreads) {                // 4 lines per row and plane
            const int yy = min(ya + (e >> 2), h - 1), xx = min(x0 + (e & 3) * 32, xmax);
            const float* q = R0 + (unsigned)yy * pitch + (unsigned)xx;
#pragma unroll
            for (int c = 0; c < 5; ++c) prefetch_l2(q + (size_t)c * plane);
        }
        for (int e = tid; e < (NROWS + 4) * 6; e += nthreads) {
            const int r = e / 6, l = e - r * 6;
            const int yy = min(max(ya - 2 + r, 0), h - 1), xx = min(max(x0 - 32 + l * 32, 0), xmax);
            const float* q = R1 + (unsigned)yy * pitch + (unsigned)xx;
#pragma unroll
            for (int c = 0; c < 5; ++c) prefetch_l2(q + (size_t)c * plane);
        }
    }
}

// Update tail of one warp: RG rows x 128 columns, lane = column within a 32-wide group, pixel i+1's taps in flight while
// pixel i is computed.  EDGE: the tile may stick out of the image or touch the 5-px attenuation ring.
template <bool EDGE, int RG, typename MT, typename RowFn, typename ColFn>
__device__ __forceinline__ void update_tail_pipelined(const uint4* __restrict__ R0, const uint4* __restrict__ R1,
                                                      const float2* __restrict__ F, MT* __restrict__ Mo, unsigned plane,
                                                      unsigned pitch, int w, int h, int x0, int y0, RowFn tail_row, ColFn tail_col) {
    constexpr int N = 4 * RG;
    auto issue = [&](int i, UpdTaps& t) {
        const int r = tail_row(i), cx = tail_col(i);
        const float2 f = F[r * kFbTW + cx];
        int x = x0 + cx, y = y0 + r;
        if (EDGE) { x = min(x, w - 1); y = min(y, h - 1); }
        update_issue_h(R0, R1, pitch, w, h, x, y, f.x, f.y, t);
    };
    auto finish = [&](int i, const UpdTaps& t) {
        const int r = tail_row(i), cx = tail_col(i);
        const int x = x0 + cx, y = y0 + r;
        float mm[5];
        update_finish_h<EDGE>(t, w, h, x, y, mm);
        if (!EDGE || (x < w && y < h)) store_m(Mo, plane, (unsigned)y * pitch + (unsigned)x, mm);
    };
    UpdTaps A, B;
    issue(0, A);
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        issue(i + 1, B);
        finish(i, A);
        if (i + 2 < N) issue(i + 2, A);
        finish(i + 1, B);
    }
}

// NW warps per CTA (TH % NW == 0): 8 -> 80 registers per thread at 3 CTAs/SM; 6 -> 112 registers and phase 1's 180 column
// tasks fill 94 % of the threads instead of 70 %.
template <int MH, bool RH, int TH, int NW = 8>
__global__ void __launch_bounds__(NW * 32, FastBoxCfg<MH, TH>::CTAS) k_blur_solve_box(const BlurSolveArgs a, const float reg, const bool tail_pipelined, const bool use_maps,
                                                                                 const __grid_constant__ TileMaps maps) {
    using C = FastBoxCfg<MH, TH>;
    constexpr int NT = NW * 32, RG = TH / NW;
    static_assert(TH % NW == 0 && NW <= 8, "rows must split evenly over the warps");
    extern __shared__ __align__(16) float smem[];
    float* V = smem;                                   // [5][TH][VP]
    float2* F = reinterpret_cast<float2*>(smem);       // [TH][TW], aliases V after phase 2
    float* s_red = smem + C::V_FLOATS;                 // [8][4]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFbTW, y0 = blockIdx.y * TH, p = blockIdx.z;
    const int w = a.w, h = a.h;
    const unsigned pitch = (unsigned)a.pitch, plane = (unsigned)a.plane_stride;
    using MT = typename MStore<RH>::type;
    const MT* Mp = static_cast<const MT*>(a.M) + (size_t)p * a.m_stride;
    constexpr int kLine = 128 / (int)sizeof(MT);                          // elements per 128-byte line
    BF_TRACE_STAMP(0);

    // A lone CTA of this kernel takes ~22 us (ncu, profiles/): its time is a chain of HBM round trips, not bandwidth.
    // So the whole M tile (with halo) is requested into L2 up front -- phase 1's register-window stream then pays L2
    // latency per step -- and likewise what phase 3 will read (R0 under the tile, R1 around it), which travels from
    // HBM while phases 1-2 run.
    // Rows of M and of packed R are 16-byte aligned spans (plan pitch is a multiple of 32 elements): one bulk request per
    // row.  Other layouts (stage API with odd pitches, fp32 R planes) keep the per-line requests.
    const bool span_ok = ((pitch | plane | (unsigned)a.m_stride) & (16 / (unsigned)sizeof(MT) - 1)) == 0;
    const void* R0 = nullptr;
    const void* R1 = nullptr;
    if (a.Mout) {
        R0 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p, a.nslots));
        R1 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p + 1, a.nslots));
    }
    if (use_maps) {
        // three instructions per CTA instead of ~1800 per-line requests (7 % of a CTA's lifetime in the phase trace)
        if (tid == 0) {
            prefetch_l2_box(&maps.m, x0 - C::HALO, y0 - MH, 0, p);
            if (a.Mout) {
                prefetch_l2_box(&maps.r0, 0, x0 / 32, y0, ring_slot(a.slot0, p, a.nslots));
                prefetch_l2_box(&maps.r1, 0, x0 / 32 - 1, y0 - 2, ring_slot(a.slot0, p + 1, a.nslots));
            }
        }
    } else if (span_ok) {
        constexpr int NROW = TH + 2 * MH;
        const int xs = max(x0 - C::HALO, 0), xe = min(x0 + kFbTW + C::HALO, w);
        const unsigned mbytes = ((unsigned)(xe - xs) * (unsigned)sizeof(MT) + 15u) & ~15u;
        for (int e = tid; e < 5 * NROW; e += NT) {
            const int c = e / NROW, r = e - c * NROW;
            const int yy = min(max(y0 - MH + r, 0), h - 1);
            prefetch_l2_span(Mp + (size_t)c * plane + (unsigned)yy * pitch + (unsigned)xs, mbytes);
        }
        if (RH && a.Mout) {
            const uint4* R0h = static_cast<const uint4*>(R0);
            const uint4* R1h = static_cast<const uint4*>(R1);
            const int xe0 = min(x0 + kFbTW, w), xs1 = max(x0 - 32, 0), xe1 = min(x0 + kFbTW + 32, w);
            // the last threads first: the M rows above went to the first 5 * NROW threads
            for (int e = NT - 1 - tid; e < 2 * TH + 4; e += NT) {
                if (e < TH) {
                    prefetch_l2_span(R0h + (unsigned)min(y0 + e, h - 1) * pitch + (unsigned)x0, (unsigned)(xe0 - x0) * 16u);
                } else {
                    const int yy = min(max(y0 - 2 + (e - TH), 0), h - 1);
                    prefetch_l2_span(R1h + (unsigned)yy * pitch + (unsigned)xs1, (unsigned)(xe1 - xs1) * 16u);
                }
            }
        }
    } else {
        constexpr int NL = (kFbTW + 2 * C::HALO + kLine - 1) / kLine + 1;            // lines per tile row (incl. misalignment)
        constexpr int NROW = TH + 2 * MH;
        const int xlo = max(x0 - C::HALO, 0) & ~(kLine - 1), xmaxl = max((w - 1) & ~(kLine - 1), 0);
        for (int e = tid; e < NROW * NL; e += NT) {
            const int r = e / NL, l = e - r * NL;
            const int yy = min(max(y0 - MH + r, 0), h - 1), xx = min(xlo + l * kLine, xmaxl);
            const MT* q = Mp + (unsigned)yy * pitch + (unsigned)xx;
#pragma unroll
            for (int c = 0; c < 5; ++c) prefetch_l2(q + (size_t)c * plane);
        }
    }
    if (a.Mout && !use_maps && !(RH && span_ok)) prefetch_r_block<RH, TH>(R0, R1, plane, pitch, w, h, x0, y0, tid, NT);

    BF_TRACE_STAMP(1);
    // ---------------- phase 1: vertical sums ----------------
    // 4-column groups are aligned to the image origin: a group is entirely inside the image, entirely left of it, entirely
    // right of it, or (width not a multiple of 4) the one group that straddles the right edge.  Outside columns load the nearest inside chunk and splat its edge lane when the SUM is
    // stored (replicate border; splat commutes with the sum), so the load path is branch-free.
    const bool rows_in = (y0 - MH >= 0) && (y0 + TH + MH <= h);      // block-uniform: no row clamping needed
    for (int task = tid; task < 5 * C::NC4; task += NT) {
        const int c = task / C::NC4, q = task - c * C::NC4;
        const int gx = x0 - C::HALO + 4 * q;
        const int wl = (w - 1) & ~3, kl = (w - 1) & 3;                // last group that holds a pixel, and that pixel's lane
        const int mode = gx < 0 ? 1 : (gx > wl ? 2 : ((gx == wl && kl != 3) ? 3 : 0));
        const int cgx = mode == 1 ? 0 : (mode == 2 ? wl : gx);
        const MT* src = Mp + (size_t)c * plane + (unsigned)cgx;
        float* dst = V + (size_t)c * TH * C::VP + 4 * q;
        // four instantiations: rows inside / clamped x interior column / replicated edge column (edge columns are rare
        // and were costing 7.5 FSEL per pixel when handled by selects)
        if (mode == 0 && rows_in) vertical_box_sums<MH, true, false, TH, C::PF>(src + (unsigned)(y0 - MH) * pitch, pitch, 0, 0, 0, 0, dst, C::VP);
        else vertical_box_sums<MH, false, true, TH, C::PF>(src, pitch, y0 - MH, h, mode, kl, dst, C::VP);
    }
    __syncthreads();
    BF_TRACE_STAMP(2);

    // ---------------- phase 2: horizontal sums + solve ----------------
    const int g = tid & 31, rb = tid >> 5;
    float2 fl[RG][4];
#pragma unroll
    for (int k = 0; k < RG; ++k) {
        const int r = rb + NW * k;
        float gs[5][4];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const float4* vp = reinterpret_cast<const float4*>(V + ((size_t)c * TH + r) * C::VP) + g;
            float vv[4 * C::NCH];
#pragma unroll
            for (int i = 0; i < C::NCH; ++i) {
                const float4 t = vp[i];
                vv[4 * i] = t.x; vv[4 * i + 1] = t.y; vv[4 * i + 2] = t.z; vv[4 * i + 3] = t.w;
            }
            // common part vv[D+3 .. D+2MH], summed as two interleaved chains for ILP
            float t0 = vv[C::D + 3], t1 = vv[C::D + 4];
#pragma unroll
            for (int i = C::D + 5; i + 1 <= C::D + 2 * MH; i += 2) { t0 += vv[i]; t1 += vv[i + 1]; }
            if (((2 * MH - 2) & 1) != 0) t0 += vv[C::D + 2 * MH];
            const float T = t0 + t1;
            const float l2 = vv[C::D + 2], l12 = vv[C::D + 1] + l2, l012 = vv[C::D] + l12;
            const float r1 = vv[C::D + 2 * MH + 1], r12 = r1 + vv[C::D + 2 * MH + 2], r123 = r12 + vv[C::D + 2 * MH + 3];
            gs[c][0] = T + l012;
            gs[c][1] = (T + l12) + r1;
            gs[c][2] = (T + l2) + r12;
            gs[c][3] = T + r123;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g11 = gs[0][j], g12 = gs[1][j], g22 = gs[2][j], h1 = gs[3][j], h2 = gs[4][j];
            const float det = diff_of_products(g11, g22, g12, g12) + reg;
            // det >= reg > 0 and far from the denormal range: compact plans take the 1-ulp hardware reciprocal (the IEEE
            // division costs ~9 instructions per pixel); exact plans keep the division
            float idet;
            if (RH) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(idet) : "f"(det));
            else idet = 1.f / det;
            fl[k][j].x = diff_of_products(g11, h2, g12, h1) * idet;
            fl[k][j].y = diff_of_products(g22, h1, g12, h2) * idet;
        }
    }
    __syncthreads();                                    // all reads of V done before F overwrites it
#pragma unroll
    for (int k = 0; k < RG; ++k) {
        float4* fp = reinterpret_cast<float4*>(F + (rb + NW * k) * kFbTW + 4 * g);
        fp[0] = make_float4(fl[k][0].x, fl[k][0].y, fl[k][1].x, fl[k][1].y);
        fp[1] = make_float4(fl[k][2].x, fl[k][2].y, fl[k][3].x, fl[k][3].y);
    }
    __syncthreads();
    BF_TRACE_STAMP(3);

    // ---------------- phase 3: coalesced tail ----------------
    // A warp walks DOWN one 32-pixel column group (NW = 8: 4 column groups x 2 row halves): the bottom taps of row r are
    // the top taps of row r + 1, so consecutive iterations of the same warp hit L1 (which is only 228 - 3 x 71 = 15 KB here)
    // instead of fetching every R1 line twice from L2.
    const int lane = tid & 31, wid = tid >> 5;
    constexpr bool kColWalk = (NW % 4) == 0;
    const int tail_r0 = kColWalk ? (wid >> 2) * (4 * RG) : wid * RG, tail_c0 = kColWalk ? (wid & 3) * 32 + lane : lane;
    auto tail_row = [&](int i) { return kColWalk ? tail_r0 + i : tail_r0 + (i >> 2); };
    auto tail_col = [&](int i) { return kColWalk ? tail_c0 : tail_c0 + (i & 3) * 32; };
    if (a.flow || a.Mout) {
        float2* fo = a.flow ? a.flow + (size_t)p * a.flow_stride : nullptr;
        MT* Mo = a.Mout ? static_cast<MT*>(a.Mout) + (size_t)p * a.m_stride : nullptr;
        // interior tiles (85 % at 1080p): bounds and the 5-px attenuation ring are decided once per tile
        const bool inner = (x0 >= 5) && (y0 >= 5) && (x0 + kFbTW <= w - 5) && (y0 + TH <= h - 5);
        if (RH && Mo && !fo && tail_pipelined) {
            if (inner) update_tail_pipelined<false, RG>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), F, Mo, plane, pitch, w, h, x0, y0, tail_row, tail_col);
            else update_tail_pipelined<true, RG>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), F, Mo, plane, pitch, w, h, x0, y0, tail_row, tail_col);
        } else if (inner && Mo && !fo) {
#pragma unroll 4
            for (int i = 0; i < 4 * RG; ++i) {
                const int r = tail_row(i), cx = tail_col(i);
                const int x = x0 + cx, y = y0 + r;
                const float2 f = F[r * kFbTW + cx];
                float mm[5];
                update_px_any<RH, false>(R0, R1, plane, pitch, w, h, x, y, f.x, f.y, mm);
                store_m(Mo, plane, (unsigned)y * pitch + (unsigned)x, mm);
            }
        } else
#pragma unroll 4
        for (int i = 0; i < 4 * RG; ++i) {
            const int r = tail_row(i), cx = tail_col(i);
            const int x = x0 + cx, y = y0 + r;
            if (x < w && y < h) {
                const float2 f = F[r * kFbTW + cx];
                if (fo) fo[(unsigned)y * (unsigned)a.flow_pitch + (unsigned)x] = f;
                if (Mo) {
                    float mm[5];
                    update_px_any<RH>(R0, R1, plane, pitch, w, h, x, y, f.x, f.y, mm);
                    store_m(Mo, plane, (unsigned)y * pitch + (unsigned)x, mm);
                }
            }
        }
    }
#ifdef BF_TRACE
    __syncthreads();
    BF_TRACE_STAMP(4);
#endif
    if (a.partial) {
        const float* ax = a.axes + p * 4;
        const float e00 = ax[0], e01 = ax[1], e10 = ax[2], e11 = ax[3];
        const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
        for (int roi = 0; roi < a.n_roi; ++roi) {
            const uint8_t* mk = a.masks + (size_t)roi * a.mask_stride;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
            for (int i = 0; i < 4 * RG; ++i) {
                const int r = tail_row(i), cx = tail_col(i);
                const int x = x0 + cx, y = y0 + r;
                if (x < w && y < h && mk[(size_t)y * a.mask_pitch + x] != 0) {
                    const float2 f = F[r * kFbTW + cx];
                    const float vx = f.x * e00 + f.y * e01;
                    const float vy = f.x * e10 + f.y * e11;
                    s0 += vx; s1 += vy; s2 += sqrtf(vx * vx + vy * vy); s3 += 1.f;
                }
            }
            s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
            __syncthreads();
            if (lane == 0) { s_red[wid * 4] = s0; s_red[wid * 4 + 1] = s1; s_red[wid * 4 + 2] = s2; s_red[wid * 4 + 3] = s3; }
            __syncthreads();
            if (tid < 4) {
                float t = 0.f;
#pragma unroll
                for (int i = 0; i < NW; ++i) t += s_red[i * 4 + tid];
                a.partial[(((size_t)p * a.n_roi + roi) * ncta + cta) * kRoiVals + tid] = t;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// k_blur_solve_gauss<MH>: the same fused iteration for OPTFLOW_FARNEBACK_GAUSSIAN (separable float32
// Gaussian window, sigma = 0.3*MH, replicate borders) -- config C4 runs winsize 21 (MH = 10).  No running sums here:
// every output is a (2MH+1)-tap weighted sum.  Phase 1 walks one scalar column per thread with the (2MH+1)-row window in
// registers (fully unrolled static ring); phase 2 reads 4+2MH values per channel with LDS.128 and evaluates 4 outputs;
// summation order is cv2's: centre tap first, then symmetric pairs (x[-i] + x[+i]) * ker[i].
// ---------------------------------------------------------------------------------------------------
template <int MH, int TH>
struct FastGaussCfg {
    static constexpr int HALO = (MH + 3) / 4 * 4;
    static constexpr int D = HALO - MH;
    static constexpr int NCOL = kFbTW + 2 * HALO;
    static constexpr int VP = NCOL + 4;
    static constexpr int WIN = 2 * MH + 1;
    static constexpr int NCH = (D + 2 * MH + 3) / 4 + 1;
    static constexpr int RG = TH / 8;
    static constexpr int V_FLOATS = 5 * TH * VP;
    static constexpr size_t SMEM = (size_t)(V_FLOATS + 64) * sizeof(float);
    static_assert(TH % 8 == 0, "tile height must be a multiple of 8");
    static_assert((size_t)TH * kFbTW * sizeof(float2) <= (size_t)V_FLOATS * sizeof(float), "F must fit in V");
};

template <int MH, bool RH, int TH>
__global__ void __launch_bounds__(256, TH >= 32 ? 2 : (TH >= 24 ? 3 : 4)) k_blur_solve_gauss(const BlurSolveArgs a, const WinCoef wc) {
    using C = FastGaussCfg<MH, TH>;
    using MT = typename MStore<RH>::type;
    extern __shared__ __align__(16) float smem[];
    float* V = smem;
    float2* F = reinterpret_cast<float2*>(smem);
    float* s_red = smem + C::V_FLOATS;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFbTW, y0 = blockIdx.y * TH, p = blockIdx.z;
    const int w = a.w, h = a.h;
    const unsigned pitch = (unsigned)a.pitch, plane = (unsigned)a.plane_stride;
    const MT* Mp = static_cast<const MT*>(a.M) + (size_t)p * a.m_stride;
    float ker[MH + 1];
#pragma unroll
    for (int i = 0; i <= MH; ++i) ker[i] = wc.ker[i];

    const void* R0 = nullptr;
    const void* R1 = nullptr;
    if (a.Mout) {
        R0 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p, a.nslots));
        R1 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p + 1, a.nslots));
        prefetch_r_block<RH, TH>(R0, R1, plane, pitch, w, h, x0, y0, tid, 256);
    }

    // ---------------- phase 1: vertical Gaussian, one scalar column per task ----------------
    for (int task = tid; task < 5 * C::NCOL; task += 256) {
        const int c = task / C::NCOL, col = task - c * C::NCOL;
        const int gx = min(max(x0 - C::HALO + col, 0), w - 1);
        const MT* src = Mp + (size_t)c * plane + (unsigned)gx;
        float* dst = V + (size_t)c * TH * C::VP + col;
        auto ld = [&](int i) -> float {                               // row y0 - MH + i, clamped (replicate)
            const int r = min(max(y0 - MH + i, 0), h - 1);
            return m_to_float(__ldg(src + (unsigned)r * pitch));
        };
        float win[C::WIN];
#pragma unroll
        for (int i = 0; i < C::WIN; ++i) win[i] = ld(i);
#pragma unroll
        for (int j = 0; j < TH; ++j) {
            // window of output row j: ring slots (j + k) % WIN for k = 0 .. 2MH, centre k = MH
            float sacc = win[(j + MH) % C::WIN] * ker[0];
#pragma unroll
            for (int i = 1; i <= MH; ++i) sacc += (win[(j + MH - i) % C::WIN] + win[(j + MH + i) % C::WIN]) * ker[i];
            dst[j * C::VP] = sacc;
            if (j + 1 < TH) win[j % C::WIN] = ld(j + C::WIN);         // row leaving the window is replaced by the next one
        }
    }
    __syncthreads();

    // ---------------- phase 2: horizontal Gaussian + solve ----------------
    const int g = tid & 31, rb = tid >> 5;
    float2 fl[C::RG][4];
#pragma unroll
    for (int k = 0; k < C::RG; ++k) {
        const int r = rb + 8 * k;
        float gs[5][4];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const float4* vp = reinterpret_cast<const float4*>(V + ((size_t)c * TH + r) * C::VP) + g;
            float vv[4 * C::NCH];
#pragma unroll
            for (int i = 0; i < C::NCH; ++i) {
                const float4 q4 = vp[i];
                vv[4 * i] = q4.x; vv[4 * i + 1] = q4.y; vv[4 * i + 2] = q4.z; vv[4 * i + 3] = q4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = C::HALO + j;
                float sacc = vv[ci] * ker[0];
#pragma unroll
                for (int i = 1; i <= MH; ++i) sacc += (vv[ci - i] + vv[ci + i]) * ker[i];
                gs[c][j] = sacc;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g11 = gs[0][j], g12 = gs[1][j], g22 = gs[2][j], h1 = gs[3][j], h2 = gs[4][j];
            const float det = diff_of_products(g11, g22, g12, g12) + 1e-3f;
            float idet;
            if (RH) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(idet) : "f"(det));      // see k_blur_solve_box
            else idet = 1.f / det;
            fl[k][j].x = diff_of_products(g11, h2, g12, h1) * idet;
            fl[k][j].y = diff_of_products(g22, h1, g12, h2) * idet;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < C::RG; ++k) {
        float4* fp = reinterpret_cast<float4*>(F + (rb + 8 * k) * kFbTW + 4 * g);
        fp[0] = make_float4(fl[k][0].x, fl[k][0].y, fl[k][1].x, fl[k][1].y);
        fp[1] = make_float4(fl[k][2].x, fl[k][2].y, fl[k][3].x, fl[k][3].y);
    }
    __syncthreads();

    // ---------------- phase 3: coalesced tail (same as the box kernel) ----------------
    const int lane = tid & 31, wid = tid >> 5;
    // a warp walks down one 32-pixel column group (4 column groups x 2 row halves), see k_blur_solve_box
    const int tail_r0 = (wid >> 2) * (4 * C::RG), tail_c0 = (wid & 3) * 32 + lane;
    auto tail_row = [&](int i) { return tail_r0 + i; };
    auto tail_col = [&](int) { return tail_c0; };
    if (a.flow || a.Mout) {
        float2* fo = a.flow ? a.flow + (size_t)p * a.flow_stride : nullptr;
        MT* Mo = a.Mout ? static_cast<MT*>(a.Mout) + (size_t)p * a.m_stride : nullptr;
        const bool inner = (x0 >= 5) && (y0 >= 5) && (x0 + kFbTW <= w - 5) && (y0 + TH <= h - 5);
        if (RH && Mo && !fo && w >= 2 && h >= 2) {
            if (inner) update_tail_pipelined<false, C::RG>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), F, Mo, plane, pitch, w, h, x0, y0, tail_row, tail_col);
            else update_tail_pipelined<true, C::RG>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), F, Mo, plane, pitch, w, h, x0, y0, tail_row, tail_col);
        } else
#pragma unroll 4
        for (int i = 0; i < 4 * C::RG; ++i) {
            const int r = tail_row(i), cx = tail_col(i);
            const int x = x0 + cx, y = y0 + r;
            if (x < w && y < h) {
                const float2 f = F[r * kFbTW + cx];
                if (fo) fo[(unsigned)y * (unsigned)a.flow_pitch + (unsigned)x] = f;
                if (Mo) {
                    float mm[5];
                    update_px_any<RH>(R0, R1, plane, pitch, w, h, x, y, f.x, f.y, mm);
                    store_m(Mo, plane, (unsigned)y * pitch + (unsigned)x, mm);
                }
            }
        }
    }
    if (a.partial) {
        const float* ax = a.axes + p * 4;
        const float e00 = ax[0], e01 = ax[1], e10 = ax[2], e11 = ax[3];
        const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
        for (int roi = 0; roi < a.n_roi; ++roi) {
            const uint8_t* mk = a.masks + (size_t)roi * a.mask_stride;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
            for (int i = 0; i < 4 * C::RG; ++i) {
                const int r = tail_row(i), cx = tail_col(i);
                const int x = x0 + cx, y = y0 + r;
                if (x < w && y < h && mk[(size_t)y * a.mask_pitch + x] != 0) {
                    const float2 f = F[r * kFbTW + cx];
                    const float vx = f.x * e00 + f.y * e01;
                    const float vy = f.x * e10 + f.y * e11;
                    s0 += vx; s1 += vy; s2 += sqrtf(vx * vx + vy * vy); s3 += 1.f;
                }
            }
            s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
            __syncthreads();
            if (lane == 0) { s_red[wid * 4] = s0; s_red[wid * 4 + 1] = s1; s_red[wid * 4 + 2] = s2; s_red[wid * 4 + 3] = s3; }
            __syncthreads();
            if (tid < 4) {
                float t = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) t += s_red[i * 4 + tid];
                a.partial[(((size_t)p * a.n_roi + roi) * ncta + cta) * kRoiVals + tid] = t;
            }
        }
    }
}

inline int gauss_th() {
    const char* e = getenv("BTCSFLOW_GAUSS_TH");
    const int v = e ? atoi(e) : 24;
    return (v == 16 || v == 32) ? v : 24;
}
inline bool gauss_fast_supported(const WinCoef& wc) { return wc.gauss && wc.m == 10; }
inline int gauss_fast_ncta(int w, int h) { const int th = gauss_th(); return ((w + kFbTW - 1) / kFbTW) * ((h + th - 1) / th); }
template <bool RH, int TH>
inline void launch_gauss_fast_t(const BlurSolveArgs& a, const WinCoef& wc, int np, cudaStream_t st) {
    using C = FastGaussCfg<10, TH>;
    cudaFuncSetAttribute(k_blur_solve_gauss<10, RH, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    dim3 g((a.w + kFbTW - 1) / kFbTW, (a.h + TH - 1) / TH, np);
    k_blur_solve_gauss<10, RH, TH><<<g, 256, C::SMEM, st>>>(a, wc);
}
inline void launch_gauss_fast(const BlurSolveArgs& a, const WinCoef& wc, int np, bool r_half, cudaStream_t st) {
    const int th = gauss_th();
    if (r_half) {
        if (th == 16) launch_gauss_fast_t<true, 16>(a, wc, np, st);
        else if (th == 32) launch_gauss_fast_t<true, 32>(a, wc, np, st);
        else launch_gauss_fast_t<true, 24>(a, wc, np, st);
    } else {
        if (th == 16) launch_gauss_fast_t<false, 16>(a, wc, np, st);
        else if (th == 32) launch_gauss_fast_t<false, 32>(a, wc, np, st);
        else launch_gauss_fast_t<false, 24>(a, wc, np, st);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline bool blur_solve_fast_supported(const WinCoef& wc, int pitch) {
    return !wc.gauss && wc.m == 7 && (pitch % 4) == 0;
}
// + at least one 4-column group and two rows (the clamped gather footprint); the row pitch (a multiple of 4 elements,
// checked above) covers the last group when the width is not a multiple of 4
inline bool blur_solve_fast_shape(int w, int h) { return w >= 4 && h >= 2; }
inline bool blur_solve_fast_aligned(const BlurSolveArgs& a) {
    return aligned16(a.M) && (a.plane_stride % 4) == 0 && (a.m_stride % 4) == 0 && blur_solve_fast_shape(a.w, a.h);
}
// Tile height.  Compact plans: 16 rows = 47 KB shared and 64 registers -> 4 CTAs/SM (the kernel is latency/issue-bound,
// resident warps win over the extra vertical halo -- 1.26 ms per 64-pair 1080p launch against 1.48 at 24 rows / 3 CTAs and
// 1.50 at 32 rows / 2 CTAs, profiles/); 64 registers hold only because the fp16 window of phase 1 stays packed (HalfRow).
// Exact plans (float4 window rows, 60 registers of window alone): 32 rows / 128 registers without spills is fastest
// (2.40 ms against 2.52 / 2.55 at 24 / 16 rows).  BTCSFLOW_TILE_TH overrides both.
inline int tile_th(bool r_half) {
    const char* e = getenv("BTCSFLOW_TILE_TH");
    const int v = e ? atoi(e) : (r_half ? 16 : 32);
    return (v == 16 || v == 24 || v == 32) ? v : (r_half ? 16 : 32);
}
inline int blur_solve_fast_ncta(int w, int h, bool r_half) {
    const int th = tile_th(r_half);
    return ((w + kFbTW - 1) / kFbTW) * ((h + th - 1) / th);
}

inline bool tail_pipelined() {
    const char* e = getenv("BTCSFLOW_TAIL");
    return !(e && e[0] == 's');          // "serial" selects the one-piece update tail (A/B experiments; 1.4 % slower)
}
inline int tile_warps() {
    const char* e = getenv("BTCSFLOW_TILE_WARPS");
    return (e && atoi(e) == 6) ? 6 : 8;
}
template <bool RH, int TH, int NW = 8>
inline void launch_blur_solve_fast_th(const BlurSolveArgs& a, const WinCoef& wc, int np, const TileMaps* maps, cudaStream_t st) {
    using C = FastBoxCfg<7, TH>;
    // per device/context attribute; cheap enough to set on every launch (one process may own several plans)
    cudaFuncSetAttribute(k_blur_solve_box<7, RH, TH, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    const float reg = 1e-3f / (wc.scale * wc.scale);
    dim3 g((a.w + kFbTW - 1) / kFbTW, (a.h + TH - 1) / TH, np);
    static const TileMaps none{};
    k_blur_solve_box<7, RH, TH, NW><<<g, NW * 32, C::SMEM, st>>>(a, reg, tail_pipelined(), maps != nullptr, maps ? *maps : none);
}
// maps (optional): tensor maps encoded for tile height maps_th; ignored when another tile height is selected.
template <bool RH>
inline void launch_blur_solve_fast_t(const BlurSolveArgs& a, const WinCoef& wc, int np, const TileMaps* maps, int maps_th,
                                     cudaStream_t st) {
    const int th = tile_th(RH);
    if (!RH || maps_th != th) maps = nullptr;
    switch (th) {
        case 24:
            if (tile_warps() == 6) launch_blur_solve_fast_th<RH, 24, 6>(a, wc, np, maps, st);
            else launch_blur_solve_fast_th<RH, 24>(a, wc, np, maps, st);
            break;
        case 32: launch_blur_solve_fast_th<RH, 32>(a, wc, np, maps, st); break;
        default: launch_blur_solve_fast_th<RH, 16>(a, wc, np, maps, st);
    }
}
inline void launch_blur_solve_fast(const BlurSolveArgs& a, const WinCoef& wc, int np, bool r_half, cudaStream_t st,
                                   const TileMaps* maps = nullptr, int maps_th = 0) {
    if (r_half) launch_blur_solve_fast_t<true>(a, wc, np, maps, maps_th, st);
    else launch_blur_solve_fast_t<false>(a, wc, np, nullptr, 0, st);
}

// Host side: encode the three tensor maps of one scale for tile height th.  M: fp16 planes [pair][5][h][pitch] at `M`;
// R: packed pixels [slot][h][pitch] x 16 B.  Returns false (maps unused, per-row prefetch instead) if the driver entry
// point is missing or rejects the layout.
inline bool encode_tile_maps(TileMaps* out, const void* M, const void* R, int w, int h, int pitch, size_t plane, int max_pairs,
                             int nslots, int th) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeFn>(fn);
    }
    if (!encode || w < 4 || h < 2) return false;
    using C = FastBoxCfg<7, 24>;                                     // HALO / MH do not depend on the tile height
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    {
        const cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, 5, (cuuint64_t)max_pairs};
        const cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)plane * 2, (cuuint64_t)plane * 10};
        const cuuint32_t box[4] = {(cuuint32_t)(kFbTW + 2 * C::HALO), (cuuint32_t)(th + 14), 5, 1};
        if (encode(&out->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(M), dims, strides, box, ones,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    // R rows as chunks of 32 pixels (128 words = 512 B): a box row is one long burst, not a 16-byte pixel
    if (pitch % 32 != 0) return false;
    const cuuint64_t rdims[4] = {128, (cuuint64_t)(pitch / 32), (cuuint64_t)h, (cuuint64_t)nslots};
    const cuuint64_t rstrides[3] = {512, (cuuint64_t)pitch * 16, (cuuint64_t)plane * 16};
    const cuuint32_t box0[4] = {128, (cuuint32_t)(kFbTW / 32), (cuuint32_t)th, 1};
    const cuuint32_t box1[4] = {128, (cuuint32_t)(kFbTW / 32 + 2), (cuuint32_t)(th + 4), 1};
    if (encode(&out->r0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(R), rdims, rstrides, box0, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (encode(&out->r1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(R), rdims, rstrides, box1, ones,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    return true;
}

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
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x + j < w) op[j] = pack_r(o[0][j], o[1][j], o[2][j], o[3][j], o[4][j]);
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
