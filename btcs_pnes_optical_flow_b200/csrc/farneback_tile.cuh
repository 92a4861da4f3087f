// Tile kernels with a compile-time window: one fused Farneback iteration per launch.
//
// k_blur_solve_box<MH>: flow = Solve(BoxBlur_{2MH+1}(M)) fused with M' = UpdateMatrices(flow) and/or the
// body-axis projection + ROI partial sums (reference call site optical_flow.py:173, reduction
// optical_flow.py:176-187).  One CTA = 128 x TH output pixels (compact plans: TH = 16, 256 threads, 4 CTAs per SM).
//   phase 0  tensor-map prefetches (UTMAPF) bring the M tile (G planes, h planes) and the R0 / R1 blocks into L2.
//   phase 1  vertical window sums, global -> shared.  One thread per (channel, column group): the 2MH+1 row window lives
//            in registers (fully unrolled ring), 8 bytes per row -- four fp16 columns of a G plane (consumed by FHADD, no
//            conversions) or two fp32 columns of an h plane -- so the 3 x 36 + 2 x 72 = 252 tasks of a tile fill the 256
//            threads; exact first window then add-new/subtract-old with a history bounded by the TH+2MH rows of the tile:
//            no long-range cancellation, and all-zero (static) regions stay exactly zero.
//   phase 2  horizontal window sums from shared with conflict-free LDS.128 (lane stride 16 B), 4 outputs per
//            thread sharing the common partial sum (no subtraction), then the 2x2 solve with Kahan-accurate
//            determinants.  The 1/winsize^2 scale is folded into the regulariser (reg = 1e-3 * winsize^4).
//   phase 3  flow is transposed through shared memory so that lanes own consecutive pixels again: coalesced R0
//            loads, bilinear R1 gather issued one pixel ahead, M' stores; ROI sums reduced per CTA (deterministic partials).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "farneback_common.cuh"

namespace bf {

constexpr int kFbTW = 128, kFbTH = 32;

template <int MH, int TH, bool RH>
struct FastBoxCfg {
    static constexpr int HALO = (MH + 3) / 4 * 4;
    static constexpr int D = HALO - MH;                        // unused leading columns in the halo
    static constexpr int NC4 = (kFbTW + 2 * HALO) / 4;         // 4-column groups per tile row
    static constexpr int NC2 = 2 * NC4;                        // 2-column groups per tile row
    static constexpr int VP = kFbTW + 2 * HALO + 4;            // shared row pitch (floats, multiple of 4)
    static constexpr int WIN = 2 * MH + 1;
    static constexpr int NCH = (D + 2 * MH + 3) / 4 + 1;       // float4 chunks a 4-output group reads
    // compact tiles store the vertical sums as prefix sums inside each 4-column group, one plane per component (prefix_store):
    // row = 4 planes x NC4 groups
    static constexpr int VROW = RH ? 4 * NC4 : VP;             // floats per (channel, row)
    static constexpr int V_FLOATS = 5 * TH * VROW;
    static constexpr int NTASK = RH ? 3 * NC4 + 2 * NC2 : 5 * NC2;   // phase-1 column tasks per tile (exact: two fp32 columns each)
    static constexpr int PF = TH >= 32 ? 8 : 4;                // register prefetch depth in phase 1
    // CTAs per SM: what registers (window = WIN rows x 2 registers compact, x 4 exact) and shared memory allow
    static constexpr int WREG = WIN * 2;
    static constexpr int CTAS_REG = WREG <= 30 ? 4 : (WREG <= 46 ? 3 : (WREG <= 84 ? 2 : 1));
    static constexpr size_t SMEM = (size_t)(V_FLOATS + 64) * sizeof(float);
    static constexpr int CTAS_SMEM = (int)((227u * 1024u) / (SMEM + 1024u));
    static constexpr int CTAS = CTAS_REG < CTAS_SMEM ? CTAS_REG : (CTAS_SMEM < 1 ? 1 : CTAS_SMEM);
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
        const size_t cta = blockIdx.x;
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
// A whole 4-D box (tile of M with its halo: x, y, channel, pair; block of R: 4 words, x, y, ring slot) requested into L2
// by ONE instruction through a tensor map; parts of the box outside the tensor are skipped by the hardware.  (Per-line
// prefetch.global.L2 costs one L1 tag cycle per 128-byte line, ~2700 cycles of a CTA's 40 000-cycle life; per-row
// cp.async.bulk.prefetch.L2 is worse -- a uniform-operand instruction wrapped in a lane loop; profiles/r1s.)
__device__ __forceinline__ void prefetch_l2_box(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// Tensor maps of one launch of the box tile kernel (built per scale at plan creation, bf::encode_tile_maps): r0 / r1 = packed
// R ring {128 words = 32 pixels, x / 32, y, slot}, boxes 128 x 4 x TH x 1 and 128 x 6 x (TH + 4) x 1.  (The matrices need
// no map: their blocks are contiguous spans.)
struct TileMaps { CUtensorMap r0, r1; };

// ---- one row of a column group as it sits in the phase-1 register window ----------------------------------------------
// RowH4: four fp16 columns kept packed (uint2), consumed by the mixed-precision add of sm_100a (FHADD: f32 + f16 -> f32,
// exact conversion included): no separate conversions and half the window registers.  RowF2 / RowF4: fp32 columns.
struct RowH4 {
    using Elem = __half; using Sum = float4;
    static constexpr int GW = 4;
    uint2 v;
    static __device__ __forceinline__ RowH4 load(const __half* p) { return RowH4{__ldg(reinterpret_cast<const uint2*>(p))}; }
    __device__ __forceinline__ void split(unsigned short h[4]) const { split_h2(v.x, h[0], h[1]); split_h2(v.y, h[2], h[3]); }
    __device__ __forceinline__ float4 first() const {
        unsigned short h[4]; split(h);
        return make_float4(fh_add(h[0], 0.f), fh_add(h[1], 0.f), fh_add(h[2], 0.f), fh_add(h[3], 0.f));
    }
    __device__ __forceinline__ void add_to(float4& s) const {
        unsigned short h[4]; split(h);
        s.x = fh_add(h[0], s.x); s.y = fh_add(h[1], s.y); s.z = fh_add(h[2], s.z); s.w = fh_add(h[3], s.w);
    }
    static __device__ __forceinline__ void slide(float4& s, const RowH4& nv, const RowH4& ov) {     // nv - (ov - s)
        unsigned short n[4], o[4]; nv.split(n); ov.split(o);
        s.x = fh_sub(n[0], fh_sub(o[0], s.x)); s.y = fh_sub(n[1], fh_sub(o[1], s.y));
        s.z = fh_sub(n[2], fh_sub(o[2], s.z)); s.w = fh_sub(n[3], fh_sub(o[3], s.w));
    }
};
struct RowF2 {
    using Elem = float; using Sum = float2;
    static constexpr int GW = 2;
    float2 v;
    static __device__ __forceinline__ RowF2 load(const float* p) { return RowF2{__ldg(reinterpret_cast<const float2*>(p))}; }
    __device__ __forceinline__ float2 first() const { return v; }
    __device__ __forceinline__ void add_to(float2& s) const { s.x += v.x; s.y += v.y; }
    static __device__ __forceinline__ void slide(float2& s, const RowF2& nv, const RowF2& ov) {
        s.x += nv.v.x - ov.v.x; s.y += nv.v.y - ov.v.y;
    }
};
struct RowF4 {
    using Elem = float; using Sum = float4;
    static constexpr int GW = 4;
    float4 v;
    static __device__ __forceinline__ RowF4 load(const float* p) { return RowF4{__ldg(reinterpret_cast<const float4*>(p))}; }
    __device__ __forceinline__ float4 first() const { return v; }
    __device__ __forceinline__ void add_to(float4& s) const { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    static __device__ __forceinline__ void slide(float4& s, const RowF4& nv, const RowF4& ov) {
        s.x += nv.v.x - ov.v.x; s.y += nv.v.y - ov.v.y; s.z += nv.v.z - ov.v.z; s.w += nv.v.w - ov.v.w;
    }
};

// Store the window sums of one column group.  EDGE: the group may lie outside the image (mode 1 = left of it: splat lane 0
// of the first group; 2 = right of it: splat lane kl of the last group) or straddle the right edge (3: lanes above kl take
// lane kl) -- replicate border; the splat commutes with the sum, so the load path stays branch-free.
template <bool EDGE>
__device__ __forceinline__ void sum_store(float* dst, const float4& s, int mode, int kl) {
    if (!EDGE || mode == 0) {
        *reinterpret_cast<float4*>(dst) = s;
    } else {
        const float e = (mode == 1 || kl == 0) ? s.x : (kl == 1 ? s.y : (kl == 2 ? s.z : s.w));
        const bool keep = mode == 3;
        *reinterpret_cast<float4*>(dst) = make_float4(keep ? s.x : e, (keep && kl >= 1) ? s.y : e, (keep && kl >= 2) ? s.z : e, e);
    }
}
template <bool EDGE>
__device__ __forceinline__ void sum_store(float* dst, const float2& s, int mode, int kl) {
    if (!EDGE || mode == 0) {
        *reinterpret_cast<float2*>(dst) = s;
    } else {
        const float e = (mode == 1 || kl == 0) ? s.x : s.y;
        *reinterpret_cast<float2*>(dst) = make_float2(mode == 3 ? s.x : e, e);
    }
}

// Compact tiles keep the vertical sums of a row as PREFIX sums inside each 4-column group, one plane of `ng` groups per
// component, Vp[channel][row][k][group]:
//     G planes (one 4-column task per group)   k = 0..3:  c0, c0+c1, c0+c1+c2, c0+c1+c2+c3
//     h planes (two 2-column tasks per group)  k = 0..3:  c0, c0+c1, c2, c2+c3
// The horizontal pass then needs, per channel and 4 outputs, the components of the two groups its windows start and end in
// and ONE value (the group sum) of every group in between, all read with lanes on consecutive groups (conflict-free
// LDS.32): 10 shared-memory wavefronts per warp instead of the 20 of five LDS.128 on raw sums, and 11 additions per 4
// outputs instead of 23.  (The kernel is bound by LSU wavefronts: 81 % of peak, ncu profiles/r2c.)
template <bool EDGE>
__device__ __forceinline__ void prefix_store(float* dst, const float4& s0, int mode, int kl, int ng) {
    float4 s = s0;
    if (EDGE && mode != 0) {
        const float e = (mode == 1 || kl == 0) ? s.x : (kl == 1 ? s.y : (kl == 2 ? s.z : s.w));
        const bool keep = mode == 3;
        s = make_float4(keep ? s.x : e, (keep && kl >= 1) ? s.y : e, (keep && kl >= 2) ? s.z : e, e);
    }
    const float p2 = s.x + s.y, p3 = p2 + s.z;
    dst[0] = s.x; dst[ng] = p2; dst[2 * ng] = p3; dst[3 * ng] = p3 + s.w;
}
template <bool EDGE>
__device__ __forceinline__ void prefix_store(float* dst, const float2& s0, int mode, int kl, int ng) {
    float2 s = s0;
    if (EDGE && mode != 0) {
        const float e = (mode == 1 || kl == 0) ? s.x : s.y;
        s = make_float2(mode == 3 ? s.x : e, e);
    }
    dst[0] = s.x; dst[ng] = s.x + s.y;
}

// Vertical (2MH+1)-row box sums of one column group for TH consecutive output rows; register ring window, software
// prefetch PF rows ahead.  ld(i) returns row i of the tile's TH + 2MH input rows (i ascends strictly from call to call and
// is a compile-time constant after unrolling).
// PREFIX: store in the prefix-plane layout (prefix_store; `ng` groups per plane) instead of raw sums.
template <int MH, bool EDGE, int TH, int PF, typename Row, bool PREFIX = false, typename LoadFn>
__device__ __forceinline__ void vertical_box_sums(LoadFn ld, int mode, int kl, float* __restrict__ dst, int vp, int ng = 0) {
    constexpr int WIN = 2 * MH + 1, NROW = TH + 2 * MH;
    static_assert(PF <= TH - 1, "prefetch depth exceeds the rows that follow the first window");
    Row win[WIN];
#pragma unroll
    for (int i = 0; i < WIN; ++i) win[i] = ld(i);
    typename Row::Sum s = win[0].first();
#pragma unroll
    for (int i = 1; i < WIN; ++i) win[i].add_to(s);
    if (PREFIX) prefix_store<EDGE>(dst, s, mode, kl, ng); else sum_store<EDGE>(dst, s, mode, kl);
    Row pre[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) pre[i] = ld(WIN + i);
#pragma unroll
    for (int j = 1; j < TH; ++j) {
        const Row nv = pre[(j - 1) % PF];
        if (j - 1 + PF + WIN < NROW) pre[(j - 1) % PF] = ld(WIN + j - 1 + PF);
        const Row ov = win[(j - 1) % WIN];
        Row::slide(s, nv, ov);
        win[(j - 1) % WIN] = nv;
        if (PREFIX) prefix_store<EDGE>(dst + j * vp, s, mode, kl, ng); else sum_store<EDGE>(dst + j * vp, s, mode, kl);
    }
}

// Which column group a phase-1 task reads: groups are aligned to the image origin, so a group is entirely inside the image
// (mode 0), left of it (1), right of it (2), or the one group that straddles the right edge (3).
template <int GW>
__device__ __forceinline__ void column_mode(int gx, int w, int& mode, int& kl, int& cgx) {
    const int wl = (w - 1) & ~(GW - 1);                               // last group that holds a pixel
    kl = (w - 1) & (GW - 1);                                          // ... and that pixel's lane
    mode = gx < 0 ? 1 : (gx > wl ? 2 : ((gx == wl && kl != GW - 1) ? 3 : 0));
    cgx = mode == 1 ? 0 : (mode == 2 ? wl : gx);
}

// One phase-1 task on planar fp32 matrices (exact plans): rows `pitch` elements apart; interior tiles step one pointer.
template <int MH, int TH, int PF, typename Row>
__device__ __forceinline__ void column_task_planar(const float* __restrict__ plane_base, unsigned pitch, int w, int h, int gx,
                                                   int y0, bool rows_in, float* __restrict__ dst, int vp) {
    int mode, kl, cgx;
    column_mode<Row::GW>(gx, w, mode, kl, cgx);
    const float* src = plane_base + (unsigned)cgx;
    if (mode == 0 && rows_in) {
        const float* pl = src + (unsigned)(y0 - MH) * pitch;
        vertical_box_sums<MH, false, TH, PF, Row>([&](int) { const Row r = Row::load(pl); pl += pitch; return r; }, 0, 0, dst, vp);
    } else {
        vertical_box_sums<MH, true, TH, PF, Row>([&](int i) { return Row::load(src + (unsigned)min(max(y0 - MH + i, 0), h - 1) * pitch); },
                                                 mode, kl, dst, vp);
    }
}

// One phase-1 task on blocked matrices (compact plans; the tile IS block (bx, by), TH == kMbH): `choff` = byte offset of the
// channel inside a block, ES = element size.  Interior tiles read every row at a compile-time offset from three pointers
// (the blocks above, of, and below the tile); tiles at the image border clamp the row (replicate) and locate its block.
template <int MH, int PF, typename Row, int ES>
__device__ __forceinline__ void column_task_blocked(const MView<true>& Mv, unsigned choff, int w, int h, int gx, int y0,
                                                    bool rows_in, float* __restrict__ dst, int vp, int ng) {
    static_assert(MH <= kMbH, "the window must not reach past the neighbouring blocks");
    int mode, kl, cgx;
    column_mode<Row::GW>(gx, w, mode, kl, cgx);
    const unsigned coff = choff + (unsigned)(cgx & (kMbW - 1)) * ES;
    constexpr int RB = kMbW * ES;                                     // bytes per block row
    using Elem = typename Row::Elem;
    if (mode == 0 && rows_in) {
        const char* pc = Mv.block(cgx >> 7, y0 >> 4) + coff;
        const char* pa = pc - Mv.block_row_bytes();
        const char* pb = pc + Mv.block_row_bytes();
        vertical_box_sums<MH, false, kMbH, PF, Row, true>([&](int i) {
            const int r = i - MH;                                     // row relative to the tile (compile-time after unrolling)
            const char* q = r < 0 ? pa + (kMbH + r) * RB : (r < kMbH ? pc + r * RB : pb + (r - kMbH) * RB);
            return Row::load(reinterpret_cast<const Elem*>(q));
        }, 0, 0, dst, vp, ng);
    } else {
        const char* pcol = Mv.block(cgx >> 7, 0) + coff;
        const unsigned brow = Mv.block_row_bytes();
        vertical_box_sums<MH, true, kMbH, PF, Row, true>([&](int i) {
            const int r = min(max(y0 - MH + i, 0), h - 1);
            return Row::load(reinterpret_cast<const Elem*>(pcol + (size_t)(r >> 4) * brow + (unsigned)(r & 15) * RB));
        }, mode, kl, dst, vp, ng);
    }
}

// Lines of R that a block of rows [ya, ya+NROWS) x columns [x0, x0+128) will touch in its update tail: R0 under the
// block, R1 within +-2 rows / +-32 columns (larger flows simply miss).  Requested into L2 ahead of use (launches without
// tensor maps: stage API, fp32 R planes).
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
        const int xmax = max((int)pitch - 32, 0);
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

// One instruction requests a whole contiguous span into L2 through the bulk-copy unit.  p 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void prefetch_l2_span(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// M tile (with halo) into L2.  Blocked matrices: the tile's own block and the blocks above / below are contiguous 28 KB
// spans (three instructions; the rows of the neighbours this tile does not read are read by their own tiles a few CTAs
// earlier / later, so nothing extra travels from HBM).  Planar matrices: per 128-byte line.
template <int MH, int HALO, int TH>
__device__ __forceinline__ void prefetch_m_tile(const MView<true>& Mv, int w, int h, int x0, int y0, int tid, int) {
    if (tid == 0) {
        const int bx = x0 >> 7, by = y0 >> 4, nby = (h + kMbH - 1) / kMbH;
        prefetch_l2_span(Mv.block(bx, by), kMbBytes);
        if (by > 0) prefetch_l2_span(Mv.block(bx, by - 1), kMbBytes);
        if (by + 1 < nby) prefetch_l2_span(Mv.block(bx, by + 1), kMbBytes);
    }
}
template <int MH, int HALO, int TH>
__device__ __forceinline__ void prefetch_m_tile(const MView<false>& Mv, int w, int h, int x0, int y0, int tid, int nthreads) {
    constexpr int NROW = TH + 2 * MH;
    constexpr int NL = (kFbTW + 2 * HALO) / 32 + 2;                     // 128-byte lines per tile row (incl. misalignment)
    const int xlo = max(x0 - HALO, 0) & ~31, xmaxl = max((w - 1) & ~31, 0);
    for (int e = tid; e < NROW * NL; e += nthreads) {
        const int r = e / NL, l = e - r * NL;
        const int yy = min(max(y0 - MH + r, 0), h - 1), xx = min(xlo + l * 32, xmaxl);
        const float* q = Mv.p + (unsigned)yy * Mv.pitch + (unsigned)xx;
#pragma unroll
        for (int c = 0; c < 5; ++c) prefetch_l2(q + (size_t)c * Mv.plane);
    }
}

// Update tail of one warp on a compact plan: N pixels per lane walking DOWN one column (lane = column within a 32-wide
// group), pixel i+1's taps in flight while pixel i is computed.  A rolled loop over pixel pairs (the fully unrolled form was
// 16 KB of straight-line code per variant: a third of the tail's stall samples were instruction-cache misses; ncu,
// profiles/r2b) with every address carried as a pointer: R0 down the column, the M' block rows at +256 / +512 bytes.
// EDGE: the tile may stick out of the image or touch the 5-px attenuation ring.
template <bool EDGE, int N>
__device__ __forceinline__ void update_tail_pipelined(const uint4* __restrict__ R0, const uint4* __restrict__ R1,
                                                      const float2* __restrict__ F, const MView<true>& Mo, unsigned pitch,
                                                      int w, int h, int x0, int y0, int r0, int cx) {
    static_assert(N % 2 == 0, "pixels are processed in pairs");
    const int x = x0 + cx, xc = EDGE ? min(x, w - 1) : x;
    const float2* fp = F + r0 * kFbTW + cx;
    const uint4* q0 = R0 + (unsigned)(EDGE ? min(y0 + r0, h - 1) : y0 + r0) * pitch + (unsigned)xc;
    char* blk = Mo.block(x0 >> 7, y0 >> 4);
    char* pg = blk + MView<true>::g_off(r0, cx);
    char* ph = blk + MView<true>::h_off(r0, cx);
    int y = y0 + r0;
    auto issue = [&](UpdTaps& t, int yy, const float2* f, const uint4* q) {
        const float2 fl = *f;
        update_issue_h(q, R1, pitch, w, h, xc, EDGE ? min(yy, h - 1) : yy, fl.x, fl.y, t);
    };
    UpdTaps A, B;
    issue(A, y, fp, q0);
#pragma unroll 1
    for (int i = 0; i < N; i += 2) {
        const uint4* q1 = (!EDGE || y + 1 < h) ? q0 + pitch : q0;
        issue(B, y + 1, fp + kFbTW, q1);
        MOut<true> mm;
        update_finish_h<EDGE>(A, w, h, x, y, mm);
        if (!EDGE || (x < w && y < h)) MView<true>::store_at(pg, ph, mm);
        const uint4* q2 = (!EDGE || y + 2 < h) ? q1 + pitch : q1;
        if (i + 2 < N) issue(A, y + 2, fp + 2 * kFbTW, q2);
        update_finish_h<EDGE>(B, w, h, x, y + 1, mm);
        if (!EDGE || (x < w && y + 1 < h)) MView<true>::store_at(pg + kMbW * 2, ph + kMbW * 4, mm);
        fp += 2 * kFbTW; q0 = q2; pg += 2 * kMbW * 2; ph += 2 * kMbW * 4; y += 2;
    }
}

// Phase 3 shared by the box and Gaussian tile kernels: flow F[TH][128] (shared) -> flow store / M' = UpdateMatrices /
// projection + ROI partial sums.  A warp walks DOWN one 32-pixel column group (4 column groups x 2 row halves): the
// bottom taps of row r are the top taps of row r + 1, so consecutive iterations of the same warp hit L1 instead of
// fetching every R1 line twice from L2.
template <bool RH, int TH>
__device__ __forceinline__ void tile_tail(const BlurSolveArgs& a, const float2* __restrict__ F, float* __restrict__ s_red,
                                          const void* R0, const void* R1, int x0, int y0, int p, int cta, int ncta) {
    constexpr int N = TH / 2;                                  // pixels per thread: 8 warps = 4 column groups x 2 row halves
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int w = a.w, h = a.h;
    const unsigned pitch = (unsigned)a.pitch, plane = (unsigned)a.plane_stride;
    const int tail_r0 = (wid >> 2) * N, tail_c0 = (wid & 3) * 32 + lane;
    if (a.flow || a.Mout) {
        float2* fo = a.flow ? a.flow + (size_t)p * a.flow_stride : nullptr;
        const bool want_m = a.Mout != nullptr;
        const MView<RH> Mo(a.Mout, a.m_stride, p, plane, pitch, w);
        // interior tiles (85 % at 1080p): bounds and the 5-px attenuation ring are decided once per tile
        const bool inner = (x0 >= 5) && (y0 >= 5) && (x0 + kFbTW <= w - 5) && (y0 + TH <= h - 5);
        bool done = false;
        if constexpr (RH) {
            static_assert(TH == kMbH, "compact tiles are the 16-row blocks of the matrices");
            if (want_m && !fo) {
                if (inner) update_tail_pipelined<false, N>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), F, Mo, pitch, w, h, x0, y0, tail_r0, tail_c0);
                else update_tail_pipelined<true, N>(static_cast<const uint4*>(R0), static_cast<const uint4*>(R1), F, Mo, pitch, w, h, x0, y0, tail_r0, tail_c0);
                done = true;
            }
        }
        if (!done) {
#pragma unroll 2
            for (int i = 0; i < N; ++i) {
                const int r = tail_r0 + i, cx = tail_c0;
                const int x = x0 + cx, y = y0 + r;
                if (x < w && y < h) {
                    const float2 f = F[r * kFbTW + cx];
                    if (fo) fo[(unsigned)y * (unsigned)a.flow_pitch + (unsigned)x] = f;
                    if (want_m) {
                        MOut<RH> mm;
                        if (inner) update_px_any<RH, false>(R0, R1, plane, pitch, w, h, x, y, f.x, f.y, mm);
                        else update_px_any<RH, true>(R0, R1, plane, pitch, w, h, x, y, f.x, f.y, mm);
                        Mo.store(y, x, mm);
                    }
                }
            }
        }
    }
    if (a.partial) {
        const float* ax = a.axes + p * 4;
        const float e00 = ax[0], e01 = ax[1], e10 = ax[2], e11 = ax[3];
        const int x = x0 + tail_c0;
        for (int roi = 0; roi < a.n_roi; ++roi) {
            const int cls = a.roi_class ? a.roi_class[(size_t)roi * ncta + cta] : 2;        // CTA-uniform
            float* dst = a.partial + (((size_t)p * a.n_roi + roi) * ncta + cta) * kRoiVals;
            if (cls == 0) {                                                                 // no ROI pixel in this tile
                if (tid < 6) dst[tid] = 0.f;
                continue;
            }
            RoiAcc acc;
            if (cls == 1) {                                                                 // every in-image pixel counts: no mask loads
#pragma unroll 4
                for (int i = 0; i < N; ++i) {
                    if (x < w && y0 + tail_r0 + i < h) {
                        const float2 f = F[(tail_r0 + i) * kFbTW + tail_c0];
                        acc.add(f.x * e00 + f.y * e01, f.x * e10 + f.y * e11);
                    }
                }
            } else {
                const uint8_t* mk = a.masks + (size_t)roi * a.mask_stride + (size_t)(y0 + tail_r0) * a.mask_pitch + x;
                bool on[N];
#pragma unroll
                for (int i = 0; i < N; ++i) on[i] = x < w && y0 + tail_r0 + i < h && mk[(size_t)i * a.mask_pitch] != 0;   // loads first
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    if (on[i]) {
                        const float2 f = F[(tail_r0 + i) * kFbTW + tail_c0];
                        acc.add(f.x * e00 + f.y * e01, f.x * e10 + f.y * e11);
                    }
                }
            }
            roi_cta_store(acc, s_red, dst);
        }
    }
}

// Horizontal (2MH+1)-column box sums of 4 adjacent outputs from the prefix planes of one (channel, row).  `base` points at
// plane 0 of the group that holds halo column 4l (l = the thread's output group): the window of output j covers halo columns
// 4l + D + j .. 4l + D + j + 2MH.  Each window is normalised to: optional head (a suffix of group a), full groups a+1 .. b-1,
// optional tail (a prefix of group b); the full groups common to the four windows are summed once.
template <int MH, int D>
struct HWin {
    static __host__ __device__ constexpr int os(int j) { return (D + j) & 3; }
    static __host__ __device__ constexpr int oe(int j) { return (D + j + 2 * MH) & 3; }
    static __host__ __device__ constexpr int a(int j) { return ((D + j) >> 2) - (os(j) == 0 ? 1 : 0); }             // start on a group boundary: no head
    static __host__ __device__ constexpr int b(int j) { return ((D + j + 2 * MH) >> 2) + (oe(j) == 3 ? 1 : 0); }    // end on a group boundary: no tail
    static __host__ __device__ constexpr int cmax(int x, int y) { return x > y ? x : y; }
    static __host__ __device__ constexpr int cmin(int x, int y) { return x < y ? x : y; }
    static constexpr int AMIN = cmin(cmin(a(0), a(1)), cmin(a(2), a(3))), AMAX = cmax(cmax(a(0), a(1)), cmax(a(2), a(3)));
    static constexpr int BMIN = cmin(cmin(b(0), b(1)), cmin(b(2), b(3))), BMAX = cmax(cmax(b(0), b(1)), cmax(b(2), b(3)));
};

// HPAIR: h planes (components c0, c0+c1, c2, c2+c3).  Every index below is a compile-time constant after unrolling.
template <int MH, int D, bool HPAIR>
__device__ __forceinline__ void hwindow4(const float* __restrict__ base, int ng, float out[4]) {
    using W = HWin<MH, D>;
    auto comp = [&](int k, int g) { return base[k * ng + g]; };
    auto gsum = [&](int g) { return HPAIR ? comp(1, g) + comp(3, g) : comp(3, g); };                          // c0 + .. + c3
    auto pref = [&](int g, int k) {                                                                           // c0 + .. + ck, k = 0..2
        return HPAIR ? (k == 2 ? comp(1, g) + comp(2, g) : comp(k, g)) : comp(k, g);
    };
    float gs[W::BMAX - W::AMIN + 1];                              // group sums of every group a window may fully cover
#pragma unroll
    for (int g = W::AMIN + 1; g < W::BMAX; ++g) gs[g - W::AMIN] = gsum(g);
    float mid = 0.f;                                              // groups that are full for all four outputs
#pragma unroll
    for (int g = W::AMAX + 1; g < W::BMIN; ++g) mid = (g == W::AMAX + 1) ? gs[g - W::AMIN] : mid + gs[g - W::AMIN];
    constexpr bool have_mid = W::AMAX + 1 < W::BMIN;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float t = mid;
        bool have = have_mid;
#pragma unroll
        for (int g = W::AMIN + 1; g < W::BMAX; ++g) {
            const bool common = g > W::AMAX && g < W::BMIN;
            if (!common && g > W::a(j) && g < W::b(j)) { t = have ? t + gs[g - W::AMIN] : gs[g - W::AMIN]; have = true; }
        }
        if (W::os(j) != 0) { const float hd = gsum(W::a(j)) - pref(W::a(j), W::os(j) - 1); t = have ? t + hd : hd; have = true; }
        if (W::oe(j) != 3) { const float tl = pref(W::b(j), W::oe(j)); t = have ? t + tl : tl; have = true; }
        out[j] = t;
    }
}

// 2x2 solves of 4 adjacent outputs from their blurred matrices gs[channel][output]; flows to fl[4].
template <bool RH>
__device__ __forceinline__ void solve4(const float gs[5][4], float reg, float2 fl[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float g11 = gs[0][j], g12 = gs[1][j], g22 = gs[2][j], h1 = gs[3][j], h2 = gs[4][j];
        const float det = diff_of_products(g11, g22, g12, g12) + reg;
        // det >= reg > 0 and far from the denormal range: compact plans take the 1-ulp hardware reciprocal (the IEEE
        // division costs ~9 instructions per pixel); exact plans keep the division
        float idet;
        if (RH) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(idet) : "f"(det));
        else idet = 1.f / det;
        fl[j].x = diff_of_products(g11, h2, g12, h1) * idet;
        fl[j].y = diff_of_products(g22, h1, g12, h2) * idet;
    }
}

template <int MH, bool RH, int TH>
__global__ void __launch_bounds__(256, FastBoxCfg<MH, TH, RH>::CTAS) k_blur_solve_box(const BlurSolveArgs a, const float reg, const bool use_maps,
                                                                                              const __grid_constant__ TileMaps maps) {
    using C = FastBoxCfg<MH, TH, RH>;
    constexpr int NT = 256, NW = 8, RG = TH / NW;
    extern __shared__ __align__(16) float smem[];
    float* V = smem;                                   // [5][TH][VP]
    float2* F = reinterpret_cast<float2*>(smem);       // [TH][TW], aliases V after phase 2
    float* s_red = smem + C::V_FLOATS;                 // [8][8]
    const int tid = threadIdx.x;
    const int w = a.w, h = a.h;
    const int nbx = (w + kFbTW - 1) / kFbTW, nby = (h + TH - 1) / TH;
    const unsigned pitch = (unsigned)a.pitch, plane = (unsigned)a.plane_stride;
    const TilePos tp = decode_cta(blockIdx.x, nbx, nby, a.np, a.pair_group);
    const int x0 = tp.bx * kFbTW, y0 = tp.by * TH, p = tp.p;
    const MView<RH> Mv(const_cast<void*>(a.M), a.m_stride, p, plane, pitch, w);
    BF_TRACE_STAMP(0);

    // A lone CTA of this kernel takes ~22 us (ncu, profiles/): its time is a chain of HBM round trips, not bandwidth.
    // So the whole M tile (with halo) is requested into L2 up front -- phase 1's register-window stream then pays L2
    // latency per step -- and likewise what phase 3 will read (R0 under the tile, R1 around it), which travels from
    // HBM while phases 1-2 run.
    const void* R0 = nullptr;
    const void* R1 = nullptr;
    if (a.Mout) {
        R0 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p, a.nslots));
        R1 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p + 1, a.nslots));
    }
    prefetch_m_tile<MH, C::HALO, TH>(Mv, w, h, x0, y0, tid, NT);
    if (a.Mout) {
        if (RH && use_maps) {
            if (tid == 32) {                                            // another warp than the M prefetches
                prefetch_l2_box(&maps.r0, 0, x0 / 32, y0, ring_slot(a.slot0, p, a.nslots));
                prefetch_l2_box(&maps.r1, 0, x0 / 32 - 1, y0 - 2, ring_slot(a.slot0, p + 1, a.nslots));
            }
        } else {
            prefetch_r_block<RH, TH>(R0, R1, plane, pitch, w, h, x0, y0, tid, NT);
        }
    }

    BF_TRACE_STAMP(1);
    // ---------------- phase 1: vertical sums ----------------
    const bool rows_in = (y0 - MH >= 0) && (y0 + TH + MH <= h);      // block-uniform: no row clamping needed
    for (int task = tid; task < C::NTASK; task += NT) {
        if constexpr (RH) {
            if (task < 3 * C::NC4) {
                const int c = task / C::NC4, q = task - c * C::NC4;
                column_task_blocked<MH, C::PF, RowH4, 2>(Mv, (unsigned)c * kMbGBytes, w, h, x0 - C::HALO + 4 * q, y0, rows_in,
                                                         V + (size_t)c * TH * C::VROW + q, C::VROW, C::NC4);
            } else {
                const int t2 = task - 3 * C::NC4;
                const int c = t2 / C::NC2, q = t2 - c * C::NC2;
                column_task_blocked<MH, C::PF, RowF2, 4>(Mv, kMbHOff + (unsigned)c * kMbHBytes, w, h, x0 - C::HALO + 2 * q, y0, rows_in,
                                                         V + (size_t)(3 + c) * TH * C::VROW + (q & 1) * 2 * C::NC4 + (q >> 1), C::VROW, C::NC4);
            }
        } else {
            const int c = task / C::NC2, q = task - c * C::NC2;
            column_task_planar<MH, TH, C::PF, RowF2>(Mv.p + (size_t)c * plane, pitch, w, h, x0 - C::HALO + 2 * q, y0, rows_in,
                                                     V + (size_t)c * TH * C::VP + 2 * q, C::VP);
        }
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
        if constexpr (RH) {
#pragma unroll
            for (int c = 0; c < 3; ++c) hwindow4<MH, C::D, false>(V + ((size_t)c * TH + r) * C::VROW + g, C::NC4, gs[c]);
#pragma unroll
            for (int c = 3; c < 5; ++c) hwindow4<MH, C::D, true>(V + ((size_t)c * TH + r) * C::VROW + g, C::NC4, gs[c]);
        } else {
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
        }
        solve4<RH>(gs, reg, fl[k]);
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
    tile_tail<RH, TH>(a, F, s_red, R0, R1, x0, y0, p, tp.by * nbx + tp.bx, nbx * nby);
#ifdef BF_TRACE
    __syncthreads();
    BF_TRACE_STAMP(4);
#endif
}

// ---------------------------------------------------------------------------------------------------
// k_blur_solve_gauss<MH>: the same fused iteration for OPTFLOW_FARNEBACK_GAUSSIAN (separable float32
// Gaussian window, sigma = 0.3*MH, replicate borders) -- config C4 runs winsize 21 (MH = 10).  No running sums here:
// every output is a (2MH+1)-tap weighted sum, and the kernel is bound by instruction issue (ncu: the first version -- one
// scalar column per task, the 2MH+1 window rows in registers -- spent 280 + 130 instructions per pixel in the two passes, less
// than half of them arithmetic: window shifting, local-memory spills, per-element loads, conversions and addresses).  Both
// passes therefore work on PAIRS of fp32 values (fma.rn.f32x2 / mul.rn.f32x2 -> FFMA2 / FMUL2, coefficient operand broadcast)
// in "scatter" form: every loaded value is multiplied into the outputs it contributes to and then dropped, so neither pass
// keeps a window in registers, and every load / conversion / address serves two accumulators.  (A packed instruction issues
// at the cost of two scalar ones -- tools/ubench_f32x2.cu -- the gain is the overhead that disappears: 542 -> 344 instr/px.)
//   pairs      A = (G11, G12) and B = (h1, h2) of one pixel; the fifth channel (G22) is paired over two adjacent columns
//              in the vertical pass and runs scalar in the horizontal one
//   phase 1    one task = one column of a channel pair (or a column pair of G22): TH + 2MH loads, TH accumulator pairs,
//              <= (2MH+1) FFMA2 per load; results to shared memory as float2 (channel-interleaved planes A and B, plain G22)
//   phase 2    one thread = 4 adjacent outputs x TH/8 rows: LDS.128 = two columns of a channel pair, 4 accumulator pairs
//   shared     planes A/B store their two-column chunks even chunks first, odd chunks from chunk HOFF on: the lanes of
//              phase 2 (chunk stride 2) then read consecutive 16-byte chunks and the float2 stores of phase 1 stay
//              conflict-free (HOFF = 4 mod 8)
// Summation runs in row / column order (cv2: centre, then symmetric pairs): an fp32 reordering, ~1e-7 relative.
// ---------------------------------------------------------------------------------------------------
template <int MH, int TH>
struct FastGaussCfg {
    static constexpr int HALO = (MH + 3) / 4 * 4;
    static constexpr int D = HALO - MH;
    static constexpr int NCOL = kFbTW + 2 * HALO;               // multiple of 8
    static constexpr int VP = NCOL + 4;                         // row pitch of the plain (G22) plane, floats
    static constexpr int WIN = 2 * MH + 1;
    static constexpr int NCH = (D + 2 * MH + 3) / 4 + 1;        // LDS.128 per row of the plain plane
    // pair planes: chunk = 2 columns x 2 channels = 16 bytes
    static constexpr int NE = NCOL / 4;                         // even (= odd) chunks per row
    static constexpr int HOFF = NE + ((12 - NE % 8) % 8);       // first odd chunk: >= NE and = 4 (mod 8)
    static constexpr int ROWCH = HOFF + NE;                     // chunks per row
    static constexpr int DE = D & ~1, E = D - DE;               // phase 2 starts at the even column 4g + DE
    static constexpr int NCHK = (E + 2 * MH + 4 + 1) / 2;       // chunks per thread and row
    static constexpr int RG = TH / 8;
    static constexpr int PAIR_FLOATS = TH * ROWCH * 4;          // one pair plane
    static constexpr int V_FLOATS = 2 * PAIR_FLOATS + TH * VP;
    static constexpr size_t SMEM = (size_t)(V_FLOATS + 64 + 32) * sizeof(float);   // + reduction scratch + taps (edge tiles)
    static constexpr int CTAS_SMEM = (int)((227u * 1024u) / (SMEM + 1024u));
    static constexpr int CTAS_REG = 3;
    static constexpr int CTAS = CTAS_REG < CTAS_SMEM ? CTAS_REG : (CTAS_SMEM < 1 ? 1 : CTAS_SMEM);
    static constexpr int NTASK_A = (NCOL + 31) / 32 * 32;       // task ranges start on warp boundaries (no divergence)
    static constexpr int NTASK = 2 * NTASK_A + NCOL / 2;
    static_assert(MH >= 2 && MH <= 16, "half window out of range for the fast path");
    static_assert(TH % 8 == 0, "tile height must be a multiple of 8");
    static_assert(NCOL % 8 == 0 && HOFF % 8 == 4 && HOFF >= NE, "chunk layout");
    static_assert(4 * 31 + DE + 2 * NCHK <= NCOL, "phase 2 must stay inside the tile row");
    static_assert((size_t)TH * kFbTW * sizeof(float2) <= (size_t)V_FLOATS * sizeof(float), "F must fit in V");
};

// Vertical (2MH+1)-tap sums of TH consecutive outputs from TH + 2MH inputs, in pairs.  ld(i) = packed input row i of the
// window (row y0 - MH + i), st(j, v) = output row j.  Everything unrolls: the tap of input i in output j is ker[|i-MH-j|].
template <int MH, int TH, typename Ld, typename St>
__device__ __forceinline__ void gauss_vertical2(Ld ld, St st, const f32x2_t (&k2)[MH + 1]) {
    f32x2_t acc[TH];
#pragma unroll
    for (int i = 0; i < TH + 2 * MH; ++i) {
        const f32x2_t v = ld(i);
#pragma unroll
        for (int j = 0; j < TH; ++j) {
            const int d = i - MH - j;
            if (d == -MH) acc[j] = mul2(v, k2[MH]);
            else if (d > -MH && d <= MH) acc[j] = fma2(v, k2[d < 0 ? -d : d], acc[j]);
        }
        if (i >= 2 * MH) st(i - 2 * MH, acc[i - 2 * MH]);
    }
}

// Phase-1 task of a tile whose window leaves the image at the top or bottom (rows clamped = replicate border).  Kept out of
// line: inlined, the compiler hoists the 2 x (TH + 2MH) clamped row addresses of this rare path above the branch and
// spills them in EVERY thread (ncu: 11 local stores per pixel, 44 B/px of DRAM writes).
template <int MH, bool RH, int TH, bool PAIR_COLS>
__device__ __noinline__ void gauss_vertical_edge_task(const MView<RH> Mv, int h, int y0, int c0, int gx, bool lo_y, bool hi_y,
                                                      float* dst, int dst_pitch, const float* __restrict__ ker) {
    // PAIR_COLS: channel c0 at columns gx, gx + 1 (gx even; the halves picked by lo_y / hi_y); else channels c0, c0 + 1 at gx
    f32x2_t k2[MH + 1];
#pragma unroll
    for (int i = 0; i <= MH; ++i) k2[i] = pk2(ker[i], ker[i]);
    gauss_vertical2<MH, TH>(
        [&](int i) {
            const int y = min(max(y0 - MH + i, 0), h - 1);
            return PAIR_COLS ? pk2(Mv.load(c0, y, gx), Mv.load(c0, y, gx + 1)) : pk2(Mv.load(c0, y, gx), Mv.load(c0 + 1, y, gx));
        },
        [&](int j, f32x2_t v) {
            const float2 t = up2(v);
            *reinterpret_cast<float2*>(dst + j * dst_pitch) = PAIR_COLS ? make_float2(lo_y ? t.y : t.x, hi_y ? t.y : t.x) : t;
        }, k2);
}

template <int MH, bool RH, int TH>
__global__ void __launch_bounds__(256, FastGaussCfg<MH, TH>::CTAS) k_blur_solve_gauss(const BlurSolveArgs a, const WinCoef wc) {
    using C = FastGaussCfg<MH, TH>;
    extern __shared__ __align__(16) float smem[];
    float* V = smem;
    float2* F = reinterpret_cast<float2*>(smem);
    float* s_red = smem + C::V_FLOATS;
    const int tid = threadIdx.x;
    const int w = a.w, h = a.h;
    const int nbx = (w + kFbTW - 1) / kFbTW, nby = (h + TH - 1) / TH;
    const TilePos tp = decode_cta(blockIdx.x, nbx, nby, a.np, a.pair_group);
    const int x0 = tp.bx * kFbTW, y0 = tp.by * TH, p = tp.p;
    const unsigned pitch = (unsigned)a.pitch, plane = (unsigned)a.plane_stride;
    const MView<RH> Mv(const_cast<void*>(a.M), a.m_stride, p, plane, pitch, w);
    f32x2_t k2[MH + 1];
#pragma unroll
    for (int i = 0; i <= MH; ++i) k2[i] = pk2(wc.ker[i], wc.ker[i]);

    const void* R0 = nullptr;
    const void* R1 = nullptr;
    if (a.Mout) {
        R0 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p, a.nslots));
        R1 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p + 1, a.nslots));
        prefetch_r_block<RH, TH>(R0, R1, plane, pitch, w, h, x0, y0, tid, 256);
    }

    // ---------------- phase 1: vertical Gaussian on pairs ----------------
    const bool rows_in = (y0 - MH >= 0) && (y0 + TH + MH <= h);
    float* const ker_s = smem + C::V_FLOATS + 64;                  // taps for the out-of-line edge path (block-uniform branch)
    if (!rows_in) {
        if (tid <= MH) ker_s[tid] = wc.ker[tid];
        __syncthreads();
    }
    float* const VA = V;
    float* const VB = V + C::PAIR_FLOATS;
    float* const VS = V + 2 * C::PAIR_FLOATS;
    for (int task = tid; task < C::NTASK; task += 256) {
        if (task < 2 * C::NTASK_A) {
            // channel pair A (task < NTASK_A) or B of one column
            const bool isB = task >= C::NTASK_A;
            const int col = isB ? task - C::NTASK_A : task;
            if (col >= C::NCOL) continue;
            const int gx = min(max(x0 - C::HALO + col, 0), w - 1);
            const int chunk = col >> 1;
            float* dst = (isB ? VB : VA) + (((chunk >> 1) + (chunk & 1) * C::HOFF) * 4 + (col & 1) * 2);
            auto st = [&](int j, f32x2_t v) { *reinterpret_cast<f32x2_t*>(dst + j * (C::ROWCH * 4)) = v; };
            const int c0 = isB ? 3 : 0;
            if constexpr (RH) {
                static_assert(TH == kMbH && MH <= kMbH, "compact Gaussian tiles are the blocks of the matrices");
                if (rows_in) {
                    // blocked matrices: rows at compile-time offsets from three pointers (block above, this one, below)
                    const char* pc = Mv.block(gx >> 7, y0 >> 4);
                    const unsigned brow = Mv.block_row_bytes();
                    if (!isB) {
                        pc += (unsigned)(gx & 127) * 2u;
                        const char* pa = pc - brow;
                        const char* pb = pc + brow;
                        gauss_vertical2<MH, TH>([&](int i) {
                            const int r = i - MH;
                            const char* q = r < 0 ? pa + (kMbH + r) * 256 : (r < kMbH ? pc + r * 256 : pb + (r - kMbH) * 256);
                            return pk2(h2f(__ldg(reinterpret_cast<const unsigned short*>(q))),
                                       h2f(__ldg(reinterpret_cast<const unsigned short*>(q + kMbGBytes))));
                        }, st, k2);
                    } else {
                        pc += kMbHOff + (unsigned)(gx & 127) * 4u;
                        const char* pa = pc - brow;
                        const char* pb = pc + brow;
                        gauss_vertical2<MH, TH>([&](int i) {
                            const int r = i - MH;
                            const char* q = r < 0 ? pa + (kMbH + r) * 512 : (r < kMbH ? pc + r * 512 : pb + (r - kMbH) * 512);
                            return pk2(__ldg(reinterpret_cast<const float*>(q)), __ldg(reinterpret_cast<const float*>(q + kMbHBytes)));
                        }, st, k2);
                    }
                } else {
                    gauss_vertical_edge_task<MH, RH, TH, false>(Mv, h, y0, c0, gx, false, false, dst, C::ROWCH * 4, ker_s);
                }
            } else {
                if (rows_in) {
                    const float* pl = Mv.p + (size_t)c0 * plane + (unsigned)gx + (unsigned)(y0 - MH) * pitch;
                    gauss_vertical2<MH, TH>([&](int) { const f32x2_t v = pk2(__ldg(pl), __ldg(pl + plane)); pl += pitch; return v; }, st, k2);
                } else {
                    gauss_vertical_edge_task<MH, RH, TH, false>(Mv, h, y0, c0, gx, false, false, dst, C::ROWCH * 4, ker_s);
                }
            }
        } else {
            // G22 (channel 2), two adjacent columns: the pair is loaded from the even column xe; columns clamped to the image
            // (replicate) pick their half after the sums (the filter is linear)
            const int cp = task - 2 * C::NTASK_A;                   // column pair: tile columns 2cp, 2cp + 1
            const int x = x0 - C::HALO + 2 * cp;
            const int cl = min(max(x, 0), w - 1), chh = min(max(x + 1, 0), w - 1);
            const int xe = cl & ~1;
            const bool lo_y = (cl & 1) != 0, hi_y = (chh - xe) != 0;
            float* dst = VS + 2 * cp;
            auto st = [&](int j, f32x2_t v) {
                const float2 t = up2(v);
                *reinterpret_cast<float2*>(dst + j * C::VP) = make_float2(lo_y ? t.y : t.x, hi_y ? t.y : t.x);
            };
            if constexpr (RH) {
                auto cvt = [](unsigned u) { unsigned short l, hh; split_h2(u, l, hh); return pk2(h2f(l), h2f(hh)); };
                if (rows_in) {
                    const char* pc = Mv.block(xe >> 7, y0 >> 4) + 2 * kMbGBytes + (unsigned)(xe & 127) * 2u;
                    const unsigned brow = Mv.block_row_bytes();
                    const char* pa = pc - brow;
                    const char* pb = pc + brow;
                    gauss_vertical2<MH, TH>([&](int i) {
                        const int r = i - MH;
                        const char* q = r < 0 ? pa + (kMbH + r) * 256 : (r < kMbH ? pc + r * 256 : pb + (r - kMbH) * 256);
                        return cvt(__ldg(reinterpret_cast<const unsigned*>(q)));
                    }, st, k2);
                } else {
                    gauss_vertical_edge_task<MH, RH, TH, true>(Mv, h, y0, 2, xe, lo_y, hi_y, dst, C::VP, ker_s);
                }
            } else {
                if (rows_in) {
                    const float* pl = Mv.p + (size_t)2 * plane + (unsigned)xe + (unsigned)(y0 - MH) * pitch;
                    gauss_vertical2<MH, TH>([&](int) { const float2 t = __ldg(reinterpret_cast<const float2*>(pl)); pl += pitch; return pk2(t.x, t.y); }, st, k2);
                } else {
                    gauss_vertical_edge_task<MH, RH, TH, true>(Mv, h, y0, 2, xe, lo_y, hi_y, dst, C::VP, ker_s);
                }
            }
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
        for (int pl = 0; pl < 2; ++pl) {
            // thread g: columns 4g + DE .. of the row = chunks 2g + DE/2 + i; chunk parity is compile-time
            const ulonglong2* rowp = reinterpret_cast<const ulonglong2*>((pl ? VB : VA) + (size_t)r * (C::ROWCH * 4)) + g;
            f32x2_t acc[4];
#pragma unroll
            for (int i = 0; i < C::NCHK; ++i) {
                constexpr int S = C::DE / 2;
                const int ch = S + i;
                const ulonglong2 q = rowp[(ch >> 1) + (ch & 1) * C::HOFF];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const f32x2_t v = half ? q.y : q.x;
                    const int l = 2 * i + half;                      // column 4g + DE + l
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int d = l - (C::E + MH + j);
                        if (d == -MH) acc[j] = mul2(v, k2[MH]);
                        else if (d > -MH && d <= MH) acc[j] = fma2(v, k2[d < 0 ? -d : d], acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 t = up2(acc[j]);
                gs[pl ? 3 : 0][j] = t.x;
                gs[pl ? 4 : 1][j] = t.y;
            }
        }
        {
            const float4* vp = reinterpret_cast<const float4*>(VS + (size_t)r * C::VP) + g;
            float acc[4];
#pragma unroll
            for (int i = 0; i < C::NCH; ++i) {
                const float4 q4 = vp[i];
                const float vv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int d = 4 * i + e - (C::HALO + j);     // column 4g + 4i + e against output centre 4g + HALO + j
                        if (d == -MH) acc[j] = vv[e] * wc.ker[MH];
                        else if (d > -MH && d <= MH) acc[j] = fmaf(vv[e], wc.ker[d < 0 ? -d : d], acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) gs[2][j] = acc[j];
        }
        solve4<RH>(gs, 1e-3f, fl[k]);
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
    tile_tail<RH, TH>(a, F, s_red, R0, R1, x0, y0, p, tp.by * nbx + tp.bx, nbx * nby);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- dispatch ---------------------------------------------------------------------------------------------------------
// Half windows with a compile-time kernel: box 2..16 (winsize 4..33), Gaussian 2..16.  Tile height: compact plans 16 rows
// (47 KB shared, 64 registers at winsize 15 -> 4 CTAs/SM: the kernel is latency/issue-bound, resident warps win over the
// extra vertical halo) = the blocks of the compact matrices; exact plans (float4 window rows) 32 rows, Gaussian 24.
constexpr int kBoxThCompact = 16, kBoxThExact = 16, kGaussThCompact = 16, kGaussThExact = 16;
inline int box_tile_th(bool r_half) { return r_half ? kBoxThCompact : kBoxThExact; }
inline bool box_fast_supported(const WinCoef& wc, int pitch) { return !wc.gauss && wc.m >= 2 && wc.m <= 16 && (pitch % 4) == 0; }
inline bool gauss_fast_supported(const WinCoef& wc) { return wc.gauss && wc.m >= 2 && wc.m <= 16; }
// + at least one 4-column group and two rows (the clamped gather footprint); the row pitch (a multiple of 4 elements,
// checked above) covers the last group when the width is not a multiple of 4
inline bool tile_fast_shape(int w, int h) { return w >= 4 && h >= 2; }
inline bool tile_fast_aligned(const BlurSolveArgs& a) {
    return aligned16(a.M) && (a.plane_stride % 4) == 0 && (a.m_stride % 16) == 0 && tile_fast_shape(a.w, a.h);
}
inline int box_fast_ncta(int w, int h, bool r_half) {
    const int th = box_tile_th(r_half);
    return ((w + kFbTW - 1) / kFbTW) * ((h + th - 1) / th);
}
inline int gauss_tile_th(bool r_half) { return r_half ? kGaussThCompact : kGaussThExact; }
inline int gauss_fast_ncta(int w, int h, bool r_half) { const int th = gauss_tile_th(r_half); return ((w + kFbTW - 1) / kFbTW) * ((h + th - 1) / th); }

// One launcher per (half window, storage); the instantiations live in tile_inst_*.cu so that they compile in parallel.
template <int MH, bool RH>
void launch_box_mh(const BlurSolveArgs& a, float reg, int np, const TileMaps* maps, cudaStream_t st);
template <int MH, bool RH>
void launch_gauss_mh(const BlurSolveArgs& a, const WinCoef& wc, int np, cudaStream_t st);

#ifdef BF_TILE_INSTANTIATE
// true the first time a (kernel, device) pair is seen: the opt-in shared-memory size is a per-device function attribute
template <int KEY>
inline bool smem_attr_needed() {
    static unsigned long long seen = 0;                 // bit per device ordinal (< 64)
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen & bit) return false;
    seen |= bit;
    return true;
}
template <int MH, bool RH>
void launch_box_mh(const BlurSolveArgs& a, float reg, int np, const TileMaps* maps, cudaStream_t st) {
    constexpr int TH = RH ? kBoxThCompact : kBoxThExact;
    using C = FastBoxCfg<MH, TH, RH>;
    const unsigned g = (unsigned)(((a.w + kFbTW - 1) / kFbTW) * ((a.h + TH - 1) / TH)) * (unsigned)np;
    static const TileMaps none{};
    if (smem_attr_needed<MH * 4 + (RH ? 1 : 0)>())       // once per device; a failure would surface at the launch below
        cudaFuncSetAttribute(k_blur_solve_box<MH, RH, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    k_blur_solve_box<MH, RH, TH><<<g, 256, C::SMEM, st>>>(a, reg, maps != nullptr, maps ? *maps : none);
}
template <int MH, bool RH>
void launch_gauss_mh(const BlurSolveArgs& a, const WinCoef& wc, int np, cudaStream_t st) {
    constexpr int TH = RH ? kGaussThCompact : kGaussThExact;
    using C = FastGaussCfg<MH, TH>;
    if (smem_attr_needed<MH * 4 + 2 + (RH ? 1 : 0)>())
        cudaFuncSetAttribute(k_blur_solve_gauss<MH, RH, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    const unsigned g = (unsigned)(((a.w + kFbTW - 1) / kFbTW) * ((a.h + TH - 1) / TH)) * (unsigned)np;
    k_blur_solve_gauss<MH, RH, TH><<<g, 256, C::SMEM, st>>>(a, wc);
}
#define BF_INSTANTIATE_TILE_MH(MH)                                                                             \
    template void launch_box_mh<MH, true>(const BlurSolveArgs&, float, int, const TileMaps*, cudaStream_t);    \
    template void launch_box_mh<MH, false>(const BlurSolveArgs&, float, int, const TileMaps*, cudaStream_t);   \
    template void launch_gauss_mh<MH, true>(const BlurSolveArgs&, const WinCoef&, int, cudaStream_t);          \
    template void launch_gauss_mh<MH, false>(const BlurSolveArgs&, const WinCoef&, int, cudaStream_t);
#endif

#ifndef BF_TILE_INSTANTIATE
template <bool RH>
inline void launch_box_fast_t(const BlurSolveArgs& a, const WinCoef& wc, int np, const TileMaps* maps, cudaStream_t st) {
    const float reg = 1e-3f / (wc.scale * wc.scale);
    switch (wc.m) {
#define BF_CASE(MH) case MH: launch_box_mh<MH, RH>(a, reg, np, maps, st); break;
        BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8) BF_CASE(9) BF_CASE(10) BF_CASE(11)
        BF_CASE(12) BF_CASE(13) BF_CASE(14) BF_CASE(15) BF_CASE(16)
#undef BF_CASE
        default: break;
    }
}
// maps (optional): tensor maps of this scale, encoded for this window and tile height (compact plans only).
inline void launch_box_fast(const BlurSolveArgs& a, const WinCoef& wc, int np, bool r_half, cudaStream_t st, const TileMaps* maps = nullptr) {
    if (r_half) launch_box_fast_t<true>(a, wc, np, maps, st);
    else launch_box_fast_t<false>(a, wc, np, nullptr, st);
}
template <bool RH>
inline void launch_gauss_fast_t(const BlurSolveArgs& a, const WinCoef& wc, int np, cudaStream_t st) {
    switch (wc.m) {
#define BF_CASE(MH) case MH: launch_gauss_mh<MH, RH>(a, wc, np, st); break;
        BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8) BF_CASE(9) BF_CASE(10) BF_CASE(11)
        BF_CASE(12) BF_CASE(13) BF_CASE(14) BF_CASE(15) BF_CASE(16)
#undef BF_CASE
        default: break;
    }
}
inline void launch_gauss_fast(const BlurSolveArgs& a, const WinCoef& wc, int np, bool r_half, cudaStream_t st) {
    if (r_half) launch_gauss_fast_t<true>(a, wc, np, st);
    else launch_gauss_fast_t<false>(a, wc, np, st);
}

#endif  // !BF_TILE_INSTANTIATE

// Host side: encode the R tensor maps of one scale (compact plans) for tile height th.  R: packed pixels [slot][h][pitch] x
// 16 B.  Returns false (maps unused, per-line prefetch instead) if the driver entry point is missing or rejects the layout.
inline bool encode_tile_maps(TileMaps* out, const void* R, int w, int h, int pitch, size_t plane, int nslots, int th) {
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
    if (!encode || w < 4 || h < 2 || pitch % 32 != 0) return false;
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    // R rows as chunks of 32 pixels (128 words = 512 B): a box row is one long burst, not a 16-byte pixel; the innermost
    // box coordinate is always 0 (16-byte aligned, as the bulk-tensor unit requires)
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

}  // namespace bf
