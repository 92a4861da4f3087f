"""Host-side mirror of the reference's flow stage (/root/reference/optical_flow.py) over libbtcsflow.so.

Same names, argument meaning and NaN/error behaviour as the reference functions; the dense Farneback
flow, the body-axis projection and the ROI means run in hand-written sm_100a CUDA kernels through the
C-ABI in include/btcsflow.h.  There is no CPU fallback: without the built library or a CUDA device
every compute call raises.

    calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n,
                             poly_sigma, flags)        <- cv2 call at optical_flow.py:173
    compute_roi_mean_body_flow(prev_gray, gray, ex, ey, roi_mask, fb_params)   <- optical_flow.py:136-189
    run_body_axis_flow_core(video_path, inter_npz, roi_polygon_xy, out_csv)    <- optical_flow.py:195-259
    FlowPlan.flow_series(frames, ex, ey, roi_masks)    <- the frame loop optical_flow.py:218-250, batched
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import (BF_DTYPE_F32, BF_DTYPE_U8, OPTFLOW_FARNEBACK_GAUSSIAN, OPTFLOW_USE_INITIAL_FLOW, BfParams,
                   BtcsFlowError, Cv2CompatError, check)

# Same parameter set and names as the reference (optical_flow.py:48-56).
FB_PARAMS = dict(
    pyr_scale=0.5,
    levels=3,
    winsize=15,
    iterations=3,
    poly_n=5,
    poly_sigma=1.2,
    flags=0,
)

_CV2_ASSERT = ("(-215:Assertion failed) prev0.size() == next0.size() && prev0.channels() == next0.channels() "
               "&& prev0.channels() == 1 && pyrScale_ < 1 in function 'calc'")


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _current_stream_ptr(device_index: int | None = None) -> int:
    """cudaStream_t of torch's current stream when torch+CUDA is in use, else the default stream."""
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_stream(device_index).cuda_stream)
    except ImportError:
        pass
    return 0


def _params_struct(params: dict) -> BfParams:
    p = dict(FB_PARAMS)
    unknown = set(params) - set(p)
    if unknown:
        raise TypeError(f"unknown Farneback parameter(s): {sorted(unknown)}")
    p.update(params)
    return BfParams(float(p["pyr_scale"]), int(p["levels"]), int(p["winsize"]), int(p["iterations"]),
                    int(p["poly_n"]), float(p["poly_sigma"]), int(p["flags"]))


class FlowPlan:
    """Owns the device workspace for W x H frames and one Farneback parameter set (bf_plan)."""

    def __init__(self, width: int, height: int, params: dict | None = None, max_pairs: int = 8,
                 max_rois: int = 1, device: int | None = None, exact: bool = False):
        """exact=True keeps polynomial coefficients and matrices as float32 planes (needed for float32 input frames or other
        poly_n); the default for uint8 frames with poly_n 5/7 is the compact storage (16-byte coefficient pixel, fp16 G +
        fp32 h matrices: storage only, arithmetic fp32; ~1e-6 px mean against cv2)."""
        self._lib = _lib.load()
        self.params = dict(FB_PARAMS, **(params or {}))
        self.width, self.height = int(width), int(height)
        self.max_pairs, self.max_rois = int(max_pairs), int(max_rois)
        if device is None:
            device = 0
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except ImportError:
                pass
        self.device = int(device)
        handle = C.c_void_p()
        ps = _params_struct(self.params)
        self.exact = bool(exact)
        check(self._lib.bf_plan_create_ex(C.byref(ps), self.width, self.height, self.max_pairs, self.max_rois,
                                          self.device, _lib.BF_PLAN_EXACT_F32 if exact else 0, C.byref(handle)))
        self._h = handle

    # -- lifecycle -----------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.bf_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- introspection -------------------------------------------------------------------------------
    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.bf_plan_workspace_bytes(self._h))

    def scales(self) -> list[dict]:
        """Pyramid scales coarse -> fine: dict(w, h, ksize, sigma, pitch)."""
        out = []
        for i in range(self._lib.bf_plan_num_scales(self._h)):
            w, h, k, pitch = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            s = C.c_double()
            check(self._lib.bf_plan_scale_info(self._h, i, C.byref(w), C.byref(h), C.byref(k), C.byref(s),
                                               C.byref(pitch)))
            out.append(dict(w=w.value, h=h.value, ksize=k.value, sigma=s.value, pitch=pitch.value))
        return out

    @property
    def coeff_storage_bits(self) -> int:
        return int(self._lib.bf_plan_coeff_storage(self._h))

    def level_pixels(self) -> int:
        return sum(s["w"] * s["h"] for s in self.scales())

    def profile(self, enable: bool = True) -> None:
        """Bracket every stage of the flow calls with tagged CUDA events (bf_plan_profile)."""
        check(self._lib.bf_plan_profile(self._h, int(enable)))

    def profile_read(self) -> dict:
        """Wait for the recorded events and clear the record.  dict(launches, total_ms, pair_iterations) of the
        finest-scale blur+solve launches, plus `stages`: per BF_PROF_* tag (iter_update, iter_last, update, coarse,
        expand) a dict(launches, ms, pairs)."""
        n, ms, pi = C.c_int(), C.c_double(), C.c_longlong()
        check(self._lib.bf_plan_profile_read(self._h, C.byref(n), C.byref(ms), C.byref(pi)))
        out = dict(launches=n.value, total_ms=ms.value, pair_iterations=pi.value, stages={})
        for tag, name in enumerate(_lib.BF_PROF_TAGS):
            tn, tms, tp = C.c_int(), C.c_double(), C.c_longlong()
            check(self._lib.bf_plan_profile_tag(self._h, tag, C.byref(tn), C.byref(tms), C.byref(tp)))
            out["stages"][name] = dict(launches=tn.value, ms=tms.value, pairs=tp.value)
        return out

    # -- one frame pair --------------------------------------------------------------------------------
    def flow_pair(self, prev, nxt, flow=None):
        """Dense flow of one pair.  numpy in -> numpy out (host copies inside); torch CUDA in -> torch out."""
        if _is_torch(prev) or _is_torch(nxt):
            return self._flow_pair_torch(prev, nxt, flow)
        prev, nxt, dtype = _coerce_pair(prev, nxt, self.width, self.height)
        use_init = bool(int(self.params["flags"]) & OPTFLOW_USE_INITIAL_FLOW)
        if flow is None:
            if use_init:
                raise Cv2CompatError(-1, "OPTFLOW_USE_INITIAL_FLOW needs the initial flow in `flow` (float32 [H, W, 2])")
            flow = np.empty((self.height, self.width, 2), np.float32)
            target = flow
        else:
            if not (isinstance(flow, np.ndarray) and flow.dtype == np.float32
                    and flow.shape == (self.height, self.width, 2)):
                raise Cv2CompatError(-1, "flow must be a float32 array of shape (H, W, 2)")
            # in/out with OPTFLOW_USE_INITIAL_FLOW (cv2 reads the initial flow from it), otherwise only written
            target = flow if flow.flags.c_contiguous else (np.ascontiguousarray(flow) if use_init else np.empty(flow.shape, np.float32))
        check(self._lib.bf_flow_pair_host(self._h, prev.ctypes.data, nxt.ctypes.data, dtype, prev.strides[0],
                                          target.ctypes.data, _current_stream_ptr(self.device)))
        if target is not flow:
            flow[...] = target
        return flow

    def _flow_pair_torch(self, prev, nxt, flow=None):
        import torch
        if not (_is_torch(prev) and _is_torch(nxt)) or not (prev.is_cuda and nxt.is_cuda):
            raise Cv2CompatError(-1, "prev and next must both be CUDA tensors (or both numpy arrays)")
        if prev.shape != nxt.shape or prev.dim() != 2 or tuple(prev.shape) != (self.height, self.width):
            raise Cv2CompatError(-1, _CV2_ASSERT)
        if prev.dtype == torch.uint8 and nxt.dtype == torch.uint8:
            dtype = BF_DTYPE_U8
        else:
            prev, nxt, dtype = prev.float(), nxt.float(), BF_DTYPE_F32
        if prev.stride(1) != 1:
            prev = prev.contiguous()
        if nxt.stride(1) != 1 or nxt.stride(0) != prev.stride(0):
            nxt = nxt.contiguous()
            prev = prev.contiguous()
        if flow is None:
            if int(self.params["flags"]) & OPTFLOW_USE_INITIAL_FLOW:
                raise Cv2CompatError(-1, "OPTFLOW_USE_INITIAL_FLOW needs the initial flow in `flow` (float32 CUDA tensor [H, W, 2])")
            flow = torch.empty((self.height, self.width, 2), dtype=torch.float32, device=prev.device)
        elif not (flow.is_cuda and flow.dtype == torch.float32 and flow.is_contiguous()
                  and tuple(flow.shape) == (self.height, self.width, 2)):
            raise Cv2CompatError(-1, "flow must be a contiguous float32 CUDA tensor of shape (H, W, 2)")
        check(self._lib.bf_flow_pair(self._h, prev.data_ptr(), nxt.data_ptr(), dtype,
                                     prev.stride(0) * prev.element_size(), flow.data_ptr(),
                                     _current_stream_ptr(prev.device.index)))
        return flow

    # -- series ------------------------------------------------------------------------------------------
    def flow_series(self, frames, ex=None, ey=None, roi_masks=None, return_flow: bool = False):
        """ROI-mean body-axis flow for every consecutive pair of `frames` [T, H, W] uint8.

        Returns float32 [n_roi, T, 3] = (vx_body, vy_body, mag_body); row 0 and rows whose axes are not
        finite are NaN (optical_flow.py:236-245).  ex, ey: [T, 2] (default identity axes).  roi_masks:
        [n_roi, H, W] or [H, W] bool/uint8 (default: full frame).  torch CUDA `frames` -> everything stays
        on the device and the call is asynchronous on the current stream; numpy `frames` -> host buffers,
        chunked H2D overlapped with compute, returns numpy.  With return_flow also returns the dense
        flow [T-1, H, W, 2].
        """
        if _is_torch(frames):
            return self._flow_series_torch(frames, ex, ey, roi_masks, return_flow)
        frames = np.ascontiguousarray(frames)
        if frames.dtype != np.uint8 or frames.ndim != 3 or frames.shape[1:] != (self.height, self.width):
            raise Cv2CompatError(-1, f"frames must be uint8 [T, {self.height}, {self.width}]")
        T = frames.shape[0]
        ex, ey = _axes(ex, ey, T)
        masks = _masks(roi_masks, self.height, self.width)
        if masks.shape[0] > self.max_rois:
            raise Cv2CompatError(-1, f"{masks.shape[0]} ROI masks > plan max_rois={self.max_rois}")
        out = np.empty((masks.shape[0], T, 3), np.float32)
        flow = np.empty((max(T - 1, 0), self.height, self.width, 2), np.float32) if return_flow else None
        check(self._lib.bf_flow_series_host(self._h, frames.ctypes.data, T, ex.ctypes.data, ey.ctypes.data,
                                            masks.ctypes.data, masks.shape[0], out.ctypes.data,
                                            flow.ctypes.data if flow is not None else None,
                                            _current_stream_ptr(self.device)))
        if T == 1:
            out[:] = np.nan
        return (out, flow) if return_flow else out

    def flow_series_async(self, frames, ex=None, ey=None, roi_masks=None):
        """Streaming form of flow_series for host frames: queues the call and returns a PendingSeries at once;
        `.result()` blocks until that call has finished and returns the float32 [n_roi, T, 3] series (a view of pinned
        memory owned by the pending object).  Calls on one plan run in submission order, so a caller walking a list of
        recordings can submit clip k+1 before consuming clip k.  `frames` should be pinned (e.g. torch `pin_memory()`)."""
        import ctypes as C
        import torch
        frames = np.ascontiguousarray(frames)
        if frames.dtype != np.uint8 or frames.ndim != 3 or frames.shape[1:] != (self.height, self.width):
            raise Cv2CompatError(-1, f"frames must be uint8 [T, {self.height}, {self.width}]")
        T = frames.shape[0]
        ex, ey = _axes(ex, ey, T)
        masks = _masks(roi_masks, self.height, self.width)
        if masks.shape[0] > self.max_rois:
            raise Cv2CompatError(-1, f"{masks.shape[0]} ROI masks > plan max_rois={self.max_rois}")
        out_t = torch.empty((masks.shape[0], T, 3), dtype=torch.float32, pin_memory=True)
        ticket = C.c_longlong(-1)
        check(self._lib.bf_flow_series_host_async(self._h, frames.ctypes.data, T, ex.ctypes.data, ey.ctypes.data,
                                                  masks.ctypes.data, masks.shape[0], out_t.data_ptr(), None,
                                                  _current_stream_ptr(self.device), C.byref(ticket)))
        return PendingSeries(self, int(ticket.value), out_t, frames, T)

    def _flow_series_torch(self, frames, ex, ey, roi_masks, return_flow):
        import torch
        if not frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() != 3 \
                or tuple(frames.shape[1:]) != (self.height, self.width):
            raise Cv2CompatError(-1, f"frames must be a CUDA uint8 tensor [T, {self.height}, {self.width}]")
        frames = frames.contiguous()
        dev = frames.device
        T = frames.shape[0]
        ex_t, ey_t = (_axes_torch(a, T, dev, col) for a, col in ((ex, 0), (ey, 1)))
        if roi_masks is None:
            masks = torch.ones((1, self.height, self.width), dtype=torch.uint8, device=dev)
        else:
            masks = torch.as_tensor(roi_masks, device=dev)
            if masks.dim() == 2:
                masks = masks[None]
            masks = (masks != 0).to(torch.uint8).contiguous()
        if tuple(masks.shape[1:]) != (self.height, self.width) or masks.shape[0] > self.max_rois:
            raise Cv2CompatError(-1, "roi_masks must be [n_roi <= max_rois, H, W]")
        out = torch.empty((masks.shape[0], T, 3), dtype=torch.float32, device=dev)
        if T == 1:
            out.fill_(float("nan"))
        flow = torch.empty((max(T - 1, 0), self.height, self.width, 2), dtype=torch.float32, device=dev) \
            if return_flow else None
        check(self._lib.bf_flow_series(self._h, frames.data_ptr(), T, ex_t.data_ptr(), ey_t.data_ptr(),
                                       masks.data_ptr(), masks.shape[0], out.data_ptr(),
                                       flow.data_ptr() if flow is not None else None,
                                       _current_stream_ptr(dev.index)))
        # keep the temporaries alive until the stream has consumed them
        for t in (frames, ex_t, ey_t, masks):
            t.record_stream(torch.cuda.current_stream(dev))
        return (out, flow) if return_flow else out


class PendingSeries:
    """Handle of one FlowPlan.flow_series_async call (keeps the frames and the pinned result buffer alive)."""

    def __init__(self, plan, ticket, out_t, frames, T):
        self._plan, self._ticket, self._out, self._frames, self._T = plan, ticket, out_t, frames, T
        self._done = False

    def result(self) -> np.ndarray:
        if not self._done:
            check(self._plan._lib.bf_flow_series_wait(self._plan._h, self._ticket))
            self._done = True
            self._frames = None
            if self._T == 1:
                self._out.fill_(float("nan"))
        return self._out.numpy()

    def __del__(self):
        # dropped without result(): the queued copies still read the frames and write the pinned result buffer
        try:
            if not self._done and getattr(self._plan, "_h", None):
                self._plan._lib.bf_flow_series_wait(self._plan._h, self._ticket)
        except Exception:
            pass


def _coerce_pair(prev, nxt, W: int | None = None, H: int | None = None):
    """cv2's input conventions: single channel, same size; u8 stays u8, anything else -> f32."""
    prev = np.asarray(prev)
    nxt = np.asarray(nxt)
    if prev.ndim == 3 and prev.shape[2] == 1:
        prev = prev[..., 0]
    if nxt.ndim == 3 and nxt.shape[2] == 1:
        nxt = nxt[..., 0]
    if prev.ndim != 2 or nxt.ndim != 2 or prev.shape != nxt.shape:
        raise Cv2CompatError(-1, _CV2_ASSERT)
    if W is not None and prev.shape != (H, W):
        raise Cv2CompatError(-1, f"plan is for {W}x{H} frames, got {prev.shape[1]}x{prev.shape[0]}")
    if prev.dtype == np.uint8 and nxt.dtype == np.uint8:
        dtype = BF_DTYPE_U8
        prev, nxt = np.ascontiguousarray(prev), np.ascontiguousarray(nxt)
    else:
        dtype = BF_DTYPE_F32
        prev, nxt = np.ascontiguousarray(prev, np.float32), np.ascontiguousarray(nxt, np.float32)
    return prev, nxt, dtype


def _axes(ex, ey, T: int):
    def one(a, col):
        if a is None:
            a = np.zeros((T, 2), np.float64)
            a[:, col] = 1.0
            return a
        a = np.asarray(a, np.float64)
        if a.shape == (2,):
            a = np.broadcast_to(a, (T, 2))
        if a.shape != (T, 2):
            raise Cv2CompatError(-1, f"body axes must have shape ({T}, 2) or (2,)")
        return np.ascontiguousarray(a)
    return one(ex, 0), one(ey, 1)


def _axes_torch(a, T: int, dev, col: int):
    import torch
    if a is None:
        t = torch.zeros((T, 2), dtype=torch.float64, device=dev)
        t[:, col] = 1.0
        return t
    t = torch.as_tensor(a, dtype=torch.float64, device=dev)
    if tuple(t.shape) == (2,):
        t = t.expand(T, 2)
    if tuple(t.shape) != (T, 2):
        raise Cv2CompatError(-1, f"body axes must have shape ({T}, 2) or (2,)")
    return t.contiguous()


def _masks(roi_masks, H: int, W: int) -> np.ndarray:
    if roi_masks is None:
        return np.ones((1, H, W), np.uint8)
    m = np.asarray(roi_masks)
    if m.ndim == 2:
        m = m[None]
    if m.ndim != 3 or m.shape[1:] != (H, W):
        raise Cv2CompatError(-1, f"roi_masks must be [n_roi, {H}, {W}] or [{H}, {W}]")
    return np.ascontiguousarray(m != 0, np.uint8)


# ---- plan cache (one plan per shape / parameter set / device) ------------------------------------------------
_plans: dict[tuple, FlowPlan] = {}
_plans_lock = threading.Lock()


def get_plan(width: int, height: int, params: dict | None = None, max_pairs: int = 1, max_rois: int = 1,
             device: int | None = None, exact: bool = False) -> FlowPlan:
    p = dict(FB_PARAMS, **(params or {}))
    key = (width, height, max_pairs, max_rois, device, exact, tuple(sorted(p.items())))
    with _plans_lock:
        plan = _plans.get(key)
        if plan is None:
            plan = _plans[key] = FlowPlan(width, height, p, max_pairs, max_rois, device, exact)
        return plan


def clear_plans() -> None:
    with _plans_lock:
        for p in _plans.values():
            p.close()
        _plans.clear()


# ---- reference call surface --------------------------------------------------------------------------------
def calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags):
    """Drop-in for cv2.calcOpticalFlowFarneback as called at optical_flow.py:173 (`**FB_PARAMS` works).

    Returns float32 [H, W, 2] (channel 0 = dx, 1 = dy).  Raises Cv2CompatError (a ValueError) where cv2
    raises its -215 assertion (size mismatch, multi-channel input, pyr_scale >= 1)."""
    if not float(pyr_scale) < 1.0:
        raise Cv2CompatError(-1, _CV2_ASSERT)
    if _is_torch(prev):
        import torch
        H, W = int(prev.shape[0]), int(prev.shape[1])
        if prev.dim() != 2 or tuple(prev.shape) != tuple(next.shape):
            raise Cv2CompatError(-1, _CV2_ASSERT)
        device = prev.device.index
        exact = not (prev.dtype == torch.uint8 and next.dtype == torch.uint8)
    else:
        prev, next, dt = _coerce_pair(prev, next)
        H, W = prev.shape
        device = None
        exact = dt != BF_DTYPE_U8
    params = dict(pyr_scale=pyr_scale, levels=levels, winsize=winsize, iterations=iterations, poly_n=poly_n,
                  poly_sigma=poly_sigma, flags=flags)
    return get_plan(W, H, params, max_pairs=1, max_rois=1, device=device, exact=exact).flow_pair(prev, next, flow)


def build_roi_mask(H: int, W: int, roi_polygon_xy: np.ndarray) -> np.ndarray:
    """ROI polygon -> bool mask, same rasterisation as the reference (int32 truncation + cv2.fillPoly,
    optical_flow.py:88-107).  Host side, once per clip."""
    import cv2
    poly = np.asarray(roi_polygon_xy, dtype=np.int32)
    mask = np.zeros((H, W), dtype=np.uint8)
    cv2.fillPoly(mask, [poly], 1)
    return mask.astype(bool)


def bgr_to_gray(frames_bgr):
    """cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) (optical_flow.py:227) on the GPU, bit-exact for uint8.

    frames_bgr: uint8 [T, H, W, 3] or [H, W, 3]; numpy (returns numpy) or a torch CUDA tensor (stays on the device)."""
    import torch
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise BtcsFlowError(_lib.BF_E_NODEVICE, "no CUDA device: bgr_to_gray has no CPU fallback")
    is_torch = _is_torch(frames_bgr)
    x = frames_bgr if is_torch else torch.from_numpy(np.ascontiguousarray(frames_bgr)).cuda()
    if x.dtype != torch.uint8 or x.shape[-1] != 3 or x.dim() not in (3, 4):
        raise Cv2CompatError(-1, "frames must be uint8 [T, H, W, 3] or [H, W, 3]")
    one = x.dim() == 3
    x = (x[None] if one else x).contiguous()
    T, H, W, _ = x.shape
    out = torch.empty((T, H, W), dtype=torch.uint8, device=x.device)
    check(lib.bf_bgr2gray(x.data_ptr(), T, W, H, W * 3, out.data_ptr(), W, _current_stream_ptr(x.device.index)))
    out = out[0] if one else out
    return out if is_torch else out.cpu().numpy()


def skel_index_from_time(t_sec: float, time_all: np.ndarray) -> int:
    """Largest upstream index with time_all[idx] <= t_sec, clipped (optical_flow.py:122-133)."""
    idx = int(np.searchsorted(time_all, t_sec, side="right")) - 1
    return min(max(idx, 0), len(time_all) - 1)


def compute_roi_mean_body_flow(prev_gray, gray, ex, ey, roi_mask, fb_params: dict) -> tuple[float, float, float]:
    """(vx_mean, vy_mean, mag_mean) of one frame pair inside the ROI (optical_flow.py:136-189).

    Flow is computed on the full frame; the ROI only masks the means (optical_flow.py:173, 185-187).  The
    projection, magnitude and masked means are fused into the last solve kernel; only three floats return."""
    prev_gray = np.asarray(prev_gray)
    gray = np.asarray(gray)
    if prev_gray.dtype != np.uint8 or gray.dtype != np.uint8:
        # the fused path takes uint8 gray frames -- what cvtColor yields at optical_flow.py:227
        raise Cv2CompatError(-1, "compute_roi_mean_body_flow expects uint8 gray frames")
    if prev_gray.ndim != 2 or prev_gray.shape != gray.shape:
        raise Cv2CompatError(-1, _CV2_ASSERT)
    H, W = gray.shape
    plan = get_plan(W, H, fb_params, max_pairs=1, max_rois=1)
    frames = np.stack([prev_gray, gray])
    axes_x = np.asarray(ex, np.float64).reshape(2)
    axes_y = np.asarray(ey, np.float64).reshape(2)
    out = plan.flow_series(frames, axes_x, axes_y, np.asarray(roi_mask) != 0)
    vx, vy, mag = (float(v) for v in out[0, 1])
    return vx, vy, mag


def run_body_axis_flow_core(video_path: str, inter_npz: str, roi_polygon_xy: np.ndarray, out_csv: str,
                            chunk_frames: int = 64, fb_params: dict | None = None, gray_on_device: bool = True) -> None:
    """Video -> flow.csv with the reference's columns and NaN rules (optical_flow.py:195-259).

    Decode stays on the host (cv2.VideoCapture: out of scope); BGR->gray runs on the GPU
    (bit-exact with cv2.cvtColor; gray_on_device=False keeps it on the host like the reference).  Frames are handed
    to the GPU in chunks that overlap by one frame, so every row is still the pair (frame-1, frame) and `prev`
    advances even across rows whose axes are invalid (optical_flow.py:249)."""
    import cv2
    import pandas as pd

    dat = np.load(inter_npz, allow_pickle=True)
    time_all = np.asarray(dat["time_all"], dtype=float)
    fps_npz = float(dat["fps"])
    ex_all = np.asarray(dat["ex"], dtype=float)
    ey_all = np.asarray(dat["ey"], dtype=float)

    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise RuntimeError(f"VideoCapture failed: {video_path}")
    fps = cap.get(cv2.CAP_PROP_FPS)
    if fps is None or fps <= 0:
        fps = fps_npz
    fps = float(fps)
    W = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    H = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    roi_mask = build_roi_mask(H, W, roi_polygon_xy)
    params = dict(FB_PARAMS, **(fb_params or {}))
    plan = FlowPlan(W, H, params, max_pairs=min(16, max(1, chunk_frames)), max_rois=1)

    rows: list[list] = []
    carry = None  # last gray frame of the previous chunk
    frame_idx = 0
    done = False
    try:
        while not done:
            grays, meta = [], []
            while len(grays) < chunk_frames:
                ret, frame = cap.read()
                if not ret:
                    done = True
                    break
                grays.append(frame if gray_on_device else cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY))
                t_msec = cap.get(cv2.CAP_PROP_POS_MSEC)
                t_sec = float(t_msec) / 1000.0 if (t_msec is not None and t_msec > 0) else frame_idx / fps
                sk = skel_index_from_time(t_sec, time_all)
                meta.append((frame_idx, t_sec, sk))
                frame_idx += 1
            if not grays:
                break
            stack = grays if (carry is None or gray_on_device) else [carry] + grays
            off = 0 if carry is None else 1
            n_stack = len(grays) + off
            idx = [m[2] for m in meta]
            ex = np.full((n_stack, 2), np.nan)
            ey = np.full((n_stack, 2), np.nan)
            ex[off:] = ex_all[idx]
            ey[off:] = ey_all[idx]
            if gray_on_device:
                import torch
                new_gray = bgr_to_gray(torch.from_numpy(np.stack(grays)).cuda())        # [n, H, W] uint8 on the device
                dev_stack = new_gray if carry is None else torch.cat([carry[None], new_gray])
                series = plan.flow_series(dev_stack, ex, ey, roi_mask)[0].cpu().numpy()
                grays = [new_gray[-1]]                                                  # device tensor carried over
            else:
                series = plan.flow_series(np.stack(stack), ex, ey, roi_mask)[0]
            for j, (fi, t_sec, sk) in enumerate(meta):
                ok = bool(np.isfinite(ex_all[sk]).all() and np.isfinite(ey_all[sk]).all())
                vx, vy, mag = (float(v) for v in series[off + j])
                rows.append([fi, t_sec, sk, int(ok), vx, vy, mag])
            carry = grays[-1]
    finally:
        cap.release()
        plan.close()

    df = pd.DataFrame(rows, columns=["frame", "t_sec", "skel_idx", "axes_ok", "vx_body", "vy_body", "mag_body"])
    df.to_csv(out_csv, index=False)


__all__ = [
    "FB_PARAMS", "FlowPlan", "get_plan", "clear_plans", "calcOpticalFlowFarneback", "build_roi_mask",
    "skel_index_from_time", "compute_roi_mean_body_flow", "run_body_axis_flow_core", "bgr_to_gray",
    "OPTFLOW_FARNEBACK_GAUSSIAN", "OPTFLOW_USE_INITIAL_FLOW", "BtcsFlowError", "Cv2CompatError",
]
