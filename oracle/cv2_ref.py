"""The reference's own CPU implementation of the flow stage, as installed -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's hot call is ``cv2.calcOpticalFlowFarneback`` (/root/reference/optical_flow.py:173) from the
un-vendored dependency opencv-python-headless (4.13.0.92 in this image; the reference pins no version).
``/root/reference`` does not travel to the GPU box but the cv2 wheel does (same image), so this module wraps
cv2 directly and restates the ~10 lines of numpy around the call (optical_flow.py:176-189).  It is used
  * by tests as the parity oracle for the CUDA path (bit-stable run to run and across thread counts),
  * by bench.py as the `cpu_baseline` / `--impl reference` arm (kind = "reference").
The product package never imports it.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np

FB_PARAMS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)


def cv2_version() -> str:
    return cv2.__version__


def farneback(prev: np.ndarray, nxt: np.ndarray, **params) -> np.ndarray:
    p = dict(FB_PARAMS, **params)
    return cv2.calcOpticalFlowFarneback(prev, nxt, None, **p)


def roi_mean_body_flow(prev_gray, gray, ex, ey, roi_mask, fb_params) -> tuple[float, float, float]:
    """Same arithmetic as compute_roi_mean_body_flow (optical_flow.py:173-189): float32 projection,
    cv2.magnitude, three masked nanmeans."""
    flow = cv2.calcOpticalFlowFarneback(prev_gray, gray, None, **fb_params)
    u, v = flow[..., 0], flow[..., 1]
    ub = u * float(ex[0]) + v * float(ex[1])
    vb = u * float(ey[0]) + v * float(ey[1])
    mag = cv2.magnitude(ub, vb)
    m = np.asarray(roi_mask, bool)
    return float(np.nanmean(ub[m])), float(np.nanmean(vb[m])), float(np.nanmean(mag[m]))


def roi_series(frames: np.ndarray, ex, ey, roi_masks, fb_params, threads: int | None = None) -> np.ndarray:
    """[n_roi, T, 3] series with the frame-loop semantics of optical_flow.py:218-250 (row 0 NaN, NaN rows for
    non-finite axes).  Independent pairs run on a thread pool with cv2.setNumThreads(1): cv2's Farneback is
    effectively single-threaded and releases the GIL, so this is the fair multi-core
    CPU figure."""
    frames = np.asarray(frames)
    T = frames.shape[0]
    masks = np.asarray(roi_masks)
    if masks.ndim == 2:
        masks = masks[None]
    ex = np.broadcast_to(np.asarray(ex, float), (T, 2))
    ey = np.broadcast_to(np.asarray(ey, float), (T, 2))
    out = np.full((masks.shape[0], T, 3), np.nan, np.float64)
    threads = threads or os.cpu_count() or 1

    def one(t: int):
        if not (np.isfinite(ex[t]).all() and np.isfinite(ey[t]).all()):
            return t, None
        flow = cv2.calcOpticalFlowFarneback(frames[t - 1], frames[t], None, **fb_params)
        u, v = flow[..., 0], flow[..., 1]
        ub = u * float(ex[t][0]) + v * float(ex[t][1])
        vb = u * float(ey[t][0]) + v * float(ey[t][1])
        mag = cv2.magnitude(ub, vb)
        res = []
        for m in masks:
            mb = m.astype(bool)
            res.append((float(np.nanmean(ub[mb])), float(np.nanmean(vb[mb])), float(np.nanmean(mag[mb]))))
        return t, res

    prev_threads = cv2.getNumThreads()
    cv2.setNumThreads(1)
    try:
        if threads > 1:
            with ThreadPoolExecutor(threads) as pool:
                results = list(pool.map(one, range(1, T)))
        else:
            results = [one(t) for t in range(1, T)]
    finally:
        cv2.setNumThreads(prev_threads)
    for t, res in results:
        if res is not None:
            out[:, t, :] = np.asarray(res)
    return out
