/*
 * btcsflow.h -- C-ABI of libbtcsflow.so: dense Farneback optical flow -> ROI body-axis series -> PC1,
 * hand-written CUDA for sm_100a (B200).
 *
 * Drop-in boundary.  The reference (saitosatoshi-1/BTCS_PNES_optical_flow) has no FFI of its own; its hot
 * path is one third-party call plus the numpy around it.  Each entry point below names the reference
 * interface it replaces (file:line under /root/reference):
 *
 *   bf_flow_pair*        <- cv2.calcOpticalFlowFarneback(prev_gray, gray, None, **FB_PARAMS)
 *                           optical_flow.py:173 (parameter set optical_flow.py:48-56)
 *   bf_flow_series*      <- the frame loop of run_body_axis_flow_core, optical_flow.py:218-250, i.e. per
 *                           frame pair compute_roi_mean_body_flow (optical_flow.py:136-189): flow, projection
 *                           on body axes (:180-181), magnitude (:183), three ROI means (:185-187), and the
 *                           NaN rules for frame 0 / non-finite axes (:236-245)
 *   bf_pc1_sliding*      <- dynamic_pc1_sliding, optical_PCA.py:136-235
 *   bf_stage_*           <- no reference counterpart: stage-level access (pyramid level, polynomial
 *                           expansion, update-matrices, blur+solve) so every kernel can be checked against
 *                           the stage-level oracle (oracle/farneback_np.py)
 *
 * Conventions
 *   - No torch types: plain pointers and sizes.  Unless a name ends in `_host`, every data pointer is a
 *     DEVICE pointer owned by the caller, and the call is asynchronous on `stream` (a cudaStream_t passed
 *     as void*; NULL = legacy default stream).
 *   - Return value: 0 = ok; < 0 = invalid argument (mirrors cv2's assertion set, see BF_E_*); > 0 = a
 *     cudaError_t.  bf_last_error() returns a thread-local message for the last non-zero return.
 *   - A plan owns all scratch memory; it is not thread-safe; distinct plans are independent.
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef BTCSFLOW_H
#define BTCSFLOW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_OPTFLOW_USE_INITIAL_FLOW 4      /* cv2.OPTFLOW_USE_INITIAL_FLOW: bf_flow_pair* read flow_out as the initial flow */
#define BF_OPTFLOW_FARNEBACK_GAUSSIAN 256  /* cv2.OPTFLOW_FARNEBACK_GAUSSIAN */

#define BF_E_INVALID (-1)      /* bad pointer / size / parameter (cv2: error -215 assertion) */
#define BF_E_UNSUPPORTED (-2)  /* valid for cv2 but outside this library (e.g. poly_n > 16, winsize > 129) */
#define BF_E_NODEVICE (-3)     /* no usable CUDA device: hard error, never a CPU fallback */

#define BF_DTYPE_U8 0
#define BF_DTYPE_F32 1

#define BF_MAX_POLY_N 16
#define BF_MAX_WIN_HALF 64
#define BF_MAX_SCALES 16

/* The cv2.calcOpticalFlowFarneback parameter set, same names as FB_PARAMS (optical_flow.py:48-56). */
typedef struct bf_params {
    double pyr_scale;
    int levels;
    int winsize;
    int iterations;
    int poly_n;
    double poly_sigma;
    int flags;
} bf_params;

typedef struct bf_plan bf_plan;

/* ---- plan ------------------------------------------------------------------------------------------ */

/* Create a plan for W x H frames.  max_pairs = frame pairs processed per batched launch (>= 1);
 * max_rois = number of ROI masks a series call may carry (>= 0).  device = CUDA ordinal. */
int bf_plan_create(const bf_params* params, int width, int height, int max_pairs, int max_rois, int device,
                   bf_plan** out);
/* Same with flags.  BF_PLAN_EXACT_F32: keep the polynomial coefficients and the update matrices as float32 planes.  The
 * default for poly_n 5 or 7 is the compact storage: coefficients in 16 bytes per pixel (linear terms float32, quadratic
 * terms fp16) and matrices as fp16 G + float32 h formed from the rounded G -- storage only, arithmetic is fp32; ~1e-5 px
 * mean / < 1e-3 px max against cv2 -- which is range-safe for uint8 frames only.  float32 input frames require an exact
 * plan. */
#define BF_PLAN_EXACT_F32 1u
int bf_plan_create_ex(const bf_params* params, int width, int height, int max_pairs, int max_rois, int device,
                      unsigned flags, bf_plan** out);
int bf_plan_destroy(bf_plan* plan);
/* Narrowest stored polynomial coefficient in bits: 16 (compact plan) or 32 (exact plan). */
int bf_plan_coeff_storage(const bf_plan* plan);
/* Bytes of device scratch the plan owns. */
size_t bf_plan_workspace_bytes(const bf_plan* plan);
/* Scales actually used (cv2 `levels` cropped so the coarsest is >= 32 px): returns L'+1. */
int bf_plan_num_scales(const bf_plan* plan);
/* i = 0 is the coarsest scale.  Any out pointer may be NULL. */
int bf_plan_scale_info(const bf_plan* plan, int i, int* w, int* h, int* ksize, double* sigma, int* pitch);
/* Optional timing of the plan's dominant kernel (fused blur + solve [+ update] at the finest scale): when enabled,
 * every such launch is bracketed by CUDA events on the launching stream.  bf_plan_profile_read waits for the
 * recorded events, returns the number of launches, their summed device time and the summed number of frame pairs
 * they processed (one launch handles a batch of pairs for ONE iteration), then clears the record. */
int bf_plan_profile(bf_plan* plan, int enable);
int bf_plan_profile_read(bf_plan* plan, int* n_launches, double* total_ms, long long* pair_iterations);
/* While profiling is enabled every stage of the series / pair calls is bracketed the same way and tagged:
 *   BF_PROF_ITER_UPDATE  finest-scale blur + solve + UpdateMatrices launches (read M, R0, gathered R1; write M')
 *   BF_PROF_ITER_LAST    finest-scale last iteration (blur + solve + projection / ROI sums or dense flow store)
 *   BF_PROF_UPDATE       finest-scale flow upsample + first UpdateMatrices
 *   BF_PROF_COARSE       the same three at all coarser scales
 *   BF_PROF_EXPAND       pyramid + polynomial expansion of the new frames of a batch (all scales)
 * bf_plan_profile_read (above) returns ITER_UPDATE + ITER_LAST and latches the per-tag sums, which
 * bf_plan_profile_tag then reports: launches (event pairs), summed device ms, summed pairs (frames for EXPAND). */
#define BF_PROF_ITER_UPDATE 0
#define BF_PROF_ITER_LAST 1
#define BF_PROF_UPDATE 2
#define BF_PROF_COARSE 3
#define BF_PROF_EXPAND 4
#define BF_PROF_NTAGS 5
int bf_plan_profile_tag(const bf_plan* plan, int tag, int* n_launches, double* total_ms, long long* pairs);
/* Number of kernel launches issued by this library on the calling thread since the last reset. */
long long bf_launch_count(void);
void bf_launch_count_reset(void);

/* ---- flow for one frame pair (replaces cv2.calcOpticalFlowFarneback, optical_flow.py:173) ------------ */

/* prev/next: single-channel images [H, pitch_bytes], dtype BF_DTYPE_U8 or BF_DTYPE_F32.
 * flow_out: float32 [H, W, 2] C-contiguous, channel 0 = dx, 1 = dy.  With BF_OPTFLOW_USE_INITIAL_FLOW in the plan's flags it
 * is in/out like cv2's `flow` argument: on entry the initial flow (cv2 resizes it to the coarsest scale with INTER_AREA
 * and multiplies by that scale). */
int bf_flow_pair(bf_plan* plan, const void* prev, const void* next, int dtype, size_t pitch_bytes,
                 float* flow_out, void* stream);
/* Same with HOST buffers: copies both frames in, the flow out, and synchronises the stream. */
int bf_flow_pair_host(bf_plan* plan, const void* prev, const void* next, int dtype, size_t pitch_bytes,
                      float* flow_out, void* stream);

/* ---- ROI body-axis series over T frames (replaces the loop at optical_flow.py:218-250) --------------- */

/* frames: uint8 [T, H, W] contiguous.  ex, ey: float64 [T, 2] body axes per frame (row t is used for the
 * pair (t-1, t)).  roi_masks: uint8 [n_roi, H, W], non-zero = inside.  out: float32 [n_roi, T, 3] =
 * (vx_body, vy_body, mag_body); row 0 and rows with non-finite axes are NaN (optical_flow.py:236-245).
 * flow_out: optional float32 [T-1, H, W, 2] dense flow of every pair (NULL to skip: the dense field then
 * never leaves the kernels). */
int bf_flow_series(bf_plan* plan, const uint8_t* frames, int T, const double* ex, const double* ey,
                   const uint8_t* roi_masks, int n_roi, float* out, float* flow_out, void* stream);
/* Same with HOST buffers (frames, ex, ey, roi_masks, out all on the host; pinned or pageable).  Frames are
 * staged to the device in chunks on a private copy stream overlapped with compute; the result is copied
 * back and the call returns after everything has completed.  flow_out (host) optional. */
int bf_flow_series_host(bf_plan* plan, const uint8_t* frames, int T, const double* ex, const double* ey,
                        const uint8_t* roi_masks, int n_roi, float* out, float* flow_out, void* stream);
/* Streaming form of the same call (a caller walking a list of clips, optical_flow.py:195-259 called per recording): returns
 * once the work is queued and hands back a ticket; bf_flow_series_wait(ticket) blocks until that call's `out` / `flow_out`
 * are complete.  Calls on one plan execute in submission order; frames, out and flow_out must stay valid until the wait
 * and should be pinned (pageable buffers make the copies -- and hence the call -- synchronous).  ex, ey and roi_masks are
 * consumed before the call returns.  At most 8 tickets are tracked; waiting on an older one waits for a newer call. */
int bf_flow_series_host_async(bf_plan* plan, const uint8_t* frames, int T, const double* ex, const double* ey,
                              const uint8_t* roi_masks, int n_roi, float* out, float* flow_out, void* stream,
                              long long* ticket);
int bf_flow_series_wait(bf_plan* plan, long long ticket);

/* ---- sliding-window PCA -> PC1 (replaces dynamic_pc1_sliding, optical_PCA.py:136-235) ---------------- */

/* vx, vy: float64 [n] (NaN allowed).  win_n/step_n are in samples (the reference derives them from
 * win_sec/step_sec and its module-global fs, optical_PCA.py:174-175; the Python shim does the same).
 * pc1_out: float64 [n], NaN where not computable. */
int bf_pc1_sliding(const double* vx, const double* vy, int n, int win_n, int step_n, double ref_x,
                   double ref_y, int min_samples, double* pc1_out, void* stream);
/* Batched: n_series series of equal length n laid out [n_series, n]; n_cfg window configurations;
 * pc1_out: float64 [n_cfg, n_series, n].  One launch for the whole ROI x window sweep. */
int bf_pc1_sliding_batched(const double* vx, const double* vy, int n_series, int n, const int* win_n,
                           const int* step_n, int n_cfg, double ref_x, double ref_y, int min_samples,
                           double* pc1_out, void* stream);
int bf_pc1_sliding_host(const double* vx, const double* vy, int n, int win_n, int step_n, double ref_x,
                        double ref_y, int min_samples, double* pc1_out);

/* ---- BGR -> gray (replaces cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), optical_flow.py:227) ---------------- */

/* bgr: uint8 [n_frames, H, in_pitch_bytes] interleaved B,G,R (cv2.VideoCapture layout); gray: uint8 [n_frames, H,
 * out_pitch_bytes].  Bit-exact with cv2 4.x: (B*3735 + G*19235 + R*9798 + 16384) >> 15.  Device pointers. */
int bf_bgr2gray(const uint8_t* bgr, int n_frames, int width, int height, size_t in_pitch_bytes, uint8_t* gray,
                size_t out_pitch_bytes, void* stream);

/* ---- NaN-robust zero-phase band-pass (replaces bandpass_nanrobust, optical_PCA.py:96-121) ------------- */

/* x, y: float64 [n_series, n] on the device (NaN allowed).  sos: HOST float64 [n_sections, 6] in scipy layout
 * (b0 b1 b2 a0 a1 a2), e.g. scipy.signal.butter(..., output="sos") as at optical_PCA.py:64-71.  zi: HOST float64
 * [n_sections, 2] = scipy.signal.sosfilt_zi(sos), or NULL to have it computed here (bf_sosfilt_zi).  Every contiguous
 * finite run of at least 3*(2*n_sections)+1 samples is filtered like scipy.signal.sosfiltfilt(sos, run, padlen =
 * min(3*(2*n_sections), len/2 - 1)); everything else stays NaN. */
int bf_bandpass_nanrobust(const double* x, int n_series, int n, const double* sos, const double* zi, int n_sections,
                          double* y, void* stream);
/* scipy.signal.sosfilt_zi for host callers without scipy: zi [n_sections, 2]. */
int bf_sosfilt_zi(const double* sos, int n_sections, double* zi);

/* ---- stage-level entry points (kernel-by-kernel parity; device pointers) ----------------------------- */

/* Pyramid level i of the plan from one full-resolution frame -> float32 [h_i, w_i] contiguous. */
int bf_stage_level_image(bf_plan* plan, const void* frame, int dtype, size_t pitch_bytes, int scale_index,
                         float* out, void* stream);
/* Polynomial expansion of a float32 [h, w] image -> float32 [5, h, w] planes (b_y, b_x, A_yy, A_xx, A_xy). */
int bf_stage_poly_exp(const float* image, int w, int h, int poly_n, double poly_sigma, float* R_planes,
                      void* stream);
/* M = UpdateMatrices(R0, R1, flow): R0/R1/M float32 [5, h, w] planes, flow float32 [h, w, 2]. */
int bf_stage_update_matrices(const float* R0, const float* R1, const float* flow, int w, int h, float* M,
                             void* stream);
/* flow = Solve(Blur(M)): box (flags = 0) or Gaussian (flags & 256) window of `winsize`. */
int bf_stage_blur_solve(const float* M, int w, int h, int winsize, int flags, float* flow, void* stream);
/* flow_out [h, w, 2] = bilinear resize of flow_in [hs, ws, 2] times `mult` (cv2.resize INTER_LINEAR). */
int bf_stage_upsample_flow(const float* flow_in, int ws, int hs, int w, int h, float mult, float* flow_out,
                           void* stream);

const char* bf_last_error(void);
/* "major.minor" of the device the plan lives on packed as major*10+minor (100 on B200), or < 0. */
int bf_device_sm(int device);
const char* bf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BTCSFLOW_H */
