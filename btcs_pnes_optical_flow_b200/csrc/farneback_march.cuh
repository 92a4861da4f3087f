// k_flow_iter_box<MH>: one Farneback iteration, flow = Solve(BoxBlur(M)) fused with M' = UpdateMatrices(flow) and/or the
// body-axis projection + ROI sums -- strip-marching, warp-specialised version (reference call site
// /root/reference/optical_flow.py:173, reduction optical_flow.py:176-187).
//
// Why this shape (ncu, profiles/r1_blur_solve_tile.md): the tile kernel (farneback_fast.cuh) ran its three phases
// back to back inside each CTA with 2 CTAs/SM; issue slots were 34 % busy and DRAM 45 % -- latency-bound.  Here
//   * a CTA owns a 128-pixel-wide column strip and MARCHES down a segment of rows in blocks of RB = 2*(2MH+1) rows;
//   * 6 LOADER warps (one thread per (channel, float4 column)) stream the M rows exactly once -- no vertical halo
//     re-reads -- keeping the 2MH+1-row sliding window and a 6-deep prefetch queue in registers, and write vertical
//     window sums into a double-buffered shared tile; the window sum is re-summed exactly at every block start, so
//     rounding history stays bounded to one block and all-zero (static) regions stay exactly zero;
//   * 10 COMPUTE warps turn a finished block into horizontal sums (conflict-free LDS.128), the 2x2 solve, then --
//     with lanes on consecutive pixels again after a transpose through shared memory -- the R0 read, bilinear R1
//     gather, M' stores and ROI sums;
//   * loaders and compute warps meet only at named barriers (full/empty per buffer), so HBM latency of the M stream
//     overlaps the ALU-heavy horizontal pass and the gather of the previous block.
// One CTA per SM (208 KB shared, 512 threads x <=128 registers).
#pragma once
#include "farneback_fast.cuh"

namespace bf {

template <int MH>
struct MarchCfg {
    static constexpr int HALO = (MH + 3) / 4 * 4;
    static constexpr int D = HALO - MH;
    static constexpr int WIN = 2 * MH + 1;
    static constexpr int RB = 2 * WIN;                         // rows per block (multiple of WIN: static ring index)
    static constexpr int NC4 = (kFbTW + 2 * HALO) / 4;         // float4 columns per row (36)
    static constexpr int VP = kFbTW + 2 * HALO + 4;            // shared row pitch in floats (148)
    static constexpr int NCH = (D + 2 * MH + 3) / 4 + 1;
    static constexpr int PF = 6;                               // prefetch depth (RB % PF == 0)
    static constexpr int LOADER_WARPS = 6, COMPUTE_WARPS = 10;
    static constexpr int NL = LOADER_WARPS * 32, NCMP = COMPUTE_WARPS * 32, NT = NL + NCMP;
    static constexpr int GROUPS = RB * (kFbTW / 4);            // 4-pixel groups per block (960)
    static constexpr int GPT = GROUPS / NCMP;                  // groups per compute thread (3)
    static constexpr int PPT = RB * kFbTW / NCMP;              // pixels per compute thread (12)
    static constexpr int VBUF = 5 * RB * VP;                   // floats per V buffer
    static constexpr size_t SMEM = (size_t)(2 * VBUF) * sizeof(float) + (size_t)RB * kFbTW * sizeof(float2) + 64 * sizeof(float);
    static_assert(RB % PF == 0, "prefetch ring must stay phase-aligned across blocks");
    static_assert(5 * NC4 <= NL, "one loader thread per (channel, float4 column)");
    static_assert(GROUPS % NCMP == 0 && (RB * kFbTW) % NCMP == 0, "block must divide evenly over compute threads");
    static_assert(RB % COMPUTE_WARPS == 0 || (RB * 4) % COMPUTE_WARPS == 0, "phase-3 mapping");
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

constexpr int kMarchMaxRoi = 2;   // ROI accumulators live in registers across the whole march; more ROIs -> tile kernel

struct MarchArgs {
    BlurSolveArgs a;
    int seg_rows;     // output rows per segment (multiple of RB except possibly the last)
    int nseg;         // segments per strip
    int nstrip;       // strips per image
    float reg;        // 1e-3 / scale^2
};

template <int MH, bool RH>
__global__ void __launch_bounds__(MarchCfg<MH>::NT, 1) k_flow_iter_box(const MarchArgs ma) {
    using C = MarchCfg<MH>;
    extern __shared__ __align__(16) float smem[];
    float* V = smem;                                                     // [2][5][RB][VP]
    float2* F = reinterpret_cast<float2*>(smem + 2 * C::VBUF);          // [RB][TW]
    float* s_red = reinterpret_cast<float*>(F + C::RB * kFbTW);         // [COMPUTE_WARPS][4]
    const BlurSolveArgs& a = ma.a;
    const int tid = threadIdx.x;
    const int seg = blockIdx.x % ma.nseg, strip = (blockIdx.x / ma.nseg) % ma.nstrip, p = blockIdx.x / (ma.nseg * ma.nstrip);
    const int w = a.w, h = a.h;
    const unsigned pitch = (unsigned)a.pitch, plane = (unsigned)a.plane_stride;
    const int x0 = strip * kFbTW;
    const int ya = seg * ma.seg_rows, yb = min(ya + ma.seg_rows, h);
    const int nblk = (yb - ya + C::RB - 1) / C::RB;
    enum { BAR_FULL0 = 1, BAR_FULL1 = 2, BAR_EMPTY0 = 3, BAR_EMPTY1 = 4, BAR_CA = 5, BAR_CB = 6 };

    if (tid < C::NL) {
        // =========================== LOADER WARPS ===========================
        const bool active = tid < 5 * C::NC4;
        const int c = active ? tid / C::NC4 : 0, q = active ? tid - c * C::NC4 : 0;
        const int gx = x0 - C::HALO + 4 * q;
        const int mode = gx < 0 ? 1 : (gx >= w ? 2 : 0);
        const int cgx = mode == 1 ? 0 : (mode == 2 ? w - 4 : gx);
        using MT = typename MStore<RH>::type;
        const MT* src = static_cast<const MT*>(a.M) + (size_t)p * a.m_stride + (size_t)c * plane + (unsigned)cgx;
        auto ld = [&](int row) -> float4 {                               // row clamped = replicate border
            const int r = min(max(row, 0), h - 1);
            return m_load4(src + (unsigned)r * pitch);
        };
        // Far-ahead L2 prefetch (no registers held): one lane per 128-byte line pulls rows kL2Ahead ahead of the register
        // prefetch queue, so the LDG.128 stream below pays L2 latency, not HBM latency.
        constexpr int kL2Ahead = 40;
        const bool pf_lane = active && mode == 0 && ((q & (RH ? 15 : 7)) == 0);   // one lane per 128-byte line
        auto pf_row = [&](int row) {
            if (pf_lane && row < h) prefetch_l2(src + (unsigned)max(row, 0) * pitch);
        };
        float4 win[C::WIN], pre[C::PF];
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < kL2Ahead; ++i) pf_row(ya + MH + 1 + C::PF + i);
        if (active) {
#pragma unroll
            for (int i = 0; i < C::WIN; ++i) win[i] = ld(ya - MH + i);
#pragma unroll
            for (int i = 0; i < C::PF; ++i) pre[i] = ld(ya + MH + 1 + i);
        }
        for (int blk = 0; blk < nblk; ++blk) {
            const int buf = blk & 1;
            if (blk >= 2) named_bar_sync(buf ? BAR_EMPTY1 : BAR_EMPTY0, C::NT);
            if (active) {
                float* dst = V + buf * C::VBUF + (size_t)c * C::RB * C::VP + 4 * q;
                const int ybase = ya + blk * C::RB;                      // first output row of this block
#pragma unroll
                for (int j = 0; j < C::RB; ++j) {
                    // Output row n = blk*RB + j (counted from ya) uses rows ya+n-MH .. ya+n+MH.  Row ya+n+MH enters ring
                    // slot (n-1) % WIN, replacing row ya+n-MH-1; RB is a multiple of WIN and of PF, so both ring
                    // indices depend on j only and stay compile-time constants.
                    if (j > 0 || blk > 0) {
                        const int slot = (j + C::WIN - 1) % C::WIN, ps = (j + C::PF - 1) % C::PF;
                        const float4 nv = pre[ps];
                        pre[ps] = ld(ybase + j + MH + C::PF);
                        pf_row(ybase + j + MH + C::PF + kL2Ahead);
                        const float4 ov = win[slot];
                        win[slot] = nv;
                        if (j > 0) s = f4add(s, f4sub(nv, ov));
                    }
                    if (j == 0) {
                        // exact re-sum at every block start: rounding history is bounded to one block of rows
                        s = win[0];
#pragma unroll
                        for (int i = 1; i < C::WIN; ++i) s = f4add(s, win[i]);
                    }
                    float4 o = s;
                    if (mode == 1) o = make_float4(s.x, s.x, s.x, s.x);
                    else if (mode == 2) o = make_float4(s.w, s.w, s.w, s.w);
                    *reinterpret_cast<float4*>(dst + j * C::VP) = o;
                }
            }
            named_bar_arrive(buf ? BAR_FULL1 : BAR_FULL0, C::NT);
        }
        // match the compute warps' last `empty` arrivals so no barrier is left half-armed
        for (int blk = max(nblk - 2, 0); blk < nblk; ++blk) named_bar_sync((blk & 1) ? BAR_EMPTY1 : BAR_EMPTY0, C::NT);
    } else {
        // =========================== COMPUTE WARPS ===========================
        const int ctid = tid - C::NL;
        const int lane = ctid & 31, cw = ctid >> 5;
        const void* R0 = nullptr;
        const void* R1 = nullptr;
        if (a.Mout) {
            R0 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p, a.nslots));
            R1 = r_slot_ptr<RH>(a.R, a.slot_stride, ring_slot(a.slot0, p + 1, a.nslots));
        }
        float2* fo = a.flow ? a.flow + (size_t)p * a.flow_stride : nullptr;
        using MT = typename MStore<RH>::type;
        MT* Mo = a.Mout ? static_cast<MT*>(a.Mout) + (size_t)p * a.m_stride : nullptr;
        float e00 = 0.f, e01 = 0.f, e10 = 0.f, e11 = 0.f;
        if (a.partial) { const float* ax = a.axes + p * 4; e00 = ax[0]; e01 = ax[1]; e10 = ax[2]; e11 = ax[3]; }
        constexpr int MAXROI = kMarchMaxRoi;
        float acc[MAXROI][4];
#pragma unroll
        for (int r = 0; r < MAXROI; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        const int nroi = a.partial ? min(a.n_roi, MAXROI) : 0;

        for (int blk = 0; blk < nblk; ++blk) {
            const int buf = blk & 1;
            const int ybase = ya + blk * C::RB;
            // pull what this block's tail will gather into L2 while the horizontal pass runs
            if (Mo) prefetch_r_block<RH, C::RB>(R0, R1, plane, pitch, w, h, x0, ybase, ctid, C::NCMP);
            named_bar_sync(buf ? BAR_FULL1 : BAR_FULL0, C::NT);
            // ---- horizontal sums + solve: GPT groups of 4 pixels per thread ----
            float2 fl[C::GPT][4];
            const float* Vb = V + buf * C::VBUF;
#pragma unroll
            for (int k = 0; k < C::GPT; ++k) {
                const int gi = ctid + k * C::NCMP;
                const int r = gi >> 5, g = gi & 31;
                float gs[5][4];
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const float4* vp = reinterpret_cast<const float4*>(Vb + ((size_t)c * C::RB + r) * C::VP) + g;
                    float vv[4 * C::NCH];
#pragma unroll
                    for (int i = 0; i < C::NCH; ++i) {
                        const float4 t = vp[i];
                        vv[4 * i] = t.x; vv[4 * i + 1] = t.y; vv[4 * i + 2] = t.z; vv[4 * i + 3] = t.w;
                    }
                    float t0 = vv[C::D + 3], t1 = vv[C::D + 4];
#pragma unroll
                    for (int i = C::D + 5; i + 1 <= C::D + 2 * MH; i += 2) { t0 += vv[i]; t1 += vv[i + 1]; }
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
                    const float idet = 1.f / (diff_of_products(g11, g22, g12, g12) + ma.reg);
                    fl[k][j].x = diff_of_products(g11, h2, g12, h1) * idet;
                    fl[k][j].y = diff_of_products(g22, h1, g12, h2) * idet;
                }
            }
            named_bar_arrive(buf ? BAR_EMPTY1 : BAR_EMPTY0, C::NT);      // V[buf] is free again
            named_bar_sync(BAR_CA, C::NCMP);                             // previous block's tail no longer reads F
#pragma unroll
            for (int k = 0; k < C::GPT; ++k) {
                const int gi = ctid + k * C::NCMP;
                float4* fp = reinterpret_cast<float4*>(F + (gi >> 5) * kFbTW + 4 * (gi & 31));
                fp[0] = make_float4(fl[k][0].x, fl[k][0].y, fl[k][1].x, fl[k][1].y);
                fp[1] = make_float4(fl[k][2].x, fl[k][2].y, fl[k][3].x, fl[k][3].y);
            }
            named_bar_sync(BAR_CB, C::NCMP);
            // ---- coalesced tail: lanes on consecutive pixels; warp cw owns rows cw*3 .. cw*3+2 of the block ----
            if (fo || Mo) {
#pragma unroll 4
                for (int i = 0; i < C::PPT; ++i) {
                    const int r = cw * (C::RB / C::COMPUTE_WARPS) + (i >> 2), cx = (i & 3) * 32 + lane;
                    const int x = x0 + cx, y = ybase + r;
                    if (x < w && y < yb) {
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
#pragma unroll
            for (int roi = 0; roi < MAXROI; ++roi) {                     // compile-time index: acc stays in registers
                if (roi >= nroi) break;
                const uint8_t* mk = a.masks + (size_t)roi * a.mask_stride;
#pragma unroll 4
                for (int i = 0; i < C::PPT; ++i) {
                    const int r = cw * (C::RB / C::COMPUTE_WARPS) + (i >> 2), cx = (i & 3) * 32 + lane;
                    const int x = x0 + cx, y = ybase + r;
                    if (x < w && y < yb && mk[(unsigned)y * (unsigned)a.mask_pitch + (unsigned)x] != 0) {
                        const float2 f = F[r * kFbTW + cx];
                        const float vx = f.x * e00 + f.y * e01;
                        const float vy = f.x * e10 + f.y * e11;
                        acc[roi][0] += vx; acc[roi][1] += vy; acc[roi][2] += sqrtf(vx * vx + vy * vy); acc[roi][3] += 1.f;
                    }
                }
            }
        }
        if (a.partial) {
            const int ncta = ma.nseg * ma.nstrip, cta = strip * ma.nseg + seg;
#pragma unroll
            for (int roi = 0; roi < MAXROI; ++roi) {
                if (roi >= a.n_roi) break;
                float v0 = acc[roi][0], v1 = acc[roi][1], v2 = acc[roi][2], v3 = acc[roi][3];
                v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
                named_bar_sync(BAR_CA, C::NCMP);
                if (lane == 0) { s_red[cw * 4] = v0; s_red[cw * 4 + 1] = v1; s_red[cw * 4 + 2] = v2; s_red[cw * 4 + 3] = v3; }
                named_bar_sync(BAR_CB, C::NCMP);
                if (ctid < 4) {
                    float t = 0.f;
#pragma unroll
                    for (int i = 0; i < C::COMPUTE_WARPS; ++i) t += s_red[i * 4 + ctid];
                    a.partial[(((size_t)p * a.n_roi + roi) * ncta + cta) * kRoiVals + ctid] = t;
                }
            }
        }
    }
}



// Segments per strip: minimise waves x (rows per segment + halo) over 148 one-CTA SMs.
inline void march_plan(int w, int h, int np, int sm_count, int& nstrip, int& nseg, int& seg_rows) {
    using C = MarchCfg<7>;
    nstrip = (w + kFbTW - 1) / kFbTW;
    const int nblk_total = (h + C::RB - 1) / C::RB;
    long best = -1;
    nseg = 1;
    for (int s = 1; s <= nblk_total && s <= 16; ++s) {
        const int bps = (nblk_total + s - 1) / s;                       // blocks per segment
        const int segs = (nblk_total + bps - 1) / bps;
        const long items = (long)np * nstrip * segs;
        const long waves = (items + sm_count - 1) / sm_count;
        const long cost = waves * ((long)bps * C::RB + 2 * 7 + 8);      // + window fill
        if (best < 0 || cost < best) { best = cost; nseg = segs; seg_rows = bps * C::RB; }
    }
}

inline bool march_supported(const WinCoef& wc, const BlurSolveArgs& a) {
    return !wc.gauss && wc.m == 7 && blur_solve_fast_aligned(a) && (a.pitch % 4) == 0 && (a.w % 4) == 0 &&
           (!a.partial || a.n_roi <= kMarchMaxRoi);
}

inline int march_ncta(int w, int h, int np, int sm_count) {
    int nstrip, nseg, seg_rows;
    march_plan(w, h, np, sm_count, nstrip, nseg, seg_rows);
    return nstrip * nseg;
}

template <bool RH>
inline void launch_march_t(const MarchArgs& ma, int np, cudaStream_t st) {
    using C = MarchCfg<7>;
    cudaFuncSetAttribute(k_flow_iter_box<7, RH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    k_flow_iter_box<7, RH><<<np * ma.nstrip * ma.nseg, C::NT, C::SMEM, st>>>(ma);
}

inline void launch_march(const BlurSolveArgs& a, const WinCoef& wc, int np, int sm_count, bool r_half, cudaStream_t st) {
    MarchArgs ma;
    ma.a = a;
    march_plan(a.w, a.h, np, sm_count, ma.nstrip, ma.nseg, ma.seg_rows);
    ma.reg = 1e-3f / (wc.scale * wc.scale);
    if (r_half) launch_march_t<true>(ma, np, st);
    else launch_march_t<false>(ma, np, st);
}

}  // namespace bf
