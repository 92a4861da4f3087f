"""Host-side mirror of the reference's PCA stage (/root/reference/optical_PCA.py) over libbtcsflow.so.

dynamic_pc1_sliding (optical_PCA.py:136-235) runs on the GPU (bf_pc1_sliding*).  The NaN-robust zero-phase
band-pass that precedes it (optical_PCA.py:64-121) exists twice: `bandpass_nanrobust` is the reference's host version
(scipy.signal.sosfiltfilt per finite run) and `bandpass_nanrobust_device` the same arithmetic on the GPU
(bf_bandpass_nanrobust), so a flow series can go to PC1 without a host hop (`flow_to_pc1`).
The Butterworth design itself (scipy.signal.butter, once per run) stays on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import Cv2CompatError, check

# Module-level defaults with the reference's names and values (optical_PCA.py:47-58).
FLOW_CSV = "flow.csv"
OUT_CSV = "flow_pc1.csv"
fs = 30
BPF_LOW_HZ = 0.5
BPF_HIGH_HZ = 5.0
BPF_ORDER = 4
WIN_SEC = 2.0
STEP_SEC = 0.1
MIN_SAMPLES_PCA = 3


def window_samples(win_sec: float, step_sec: float, fs_hz: float) -> tuple[int, int]:
    """win_n, step_n exactly as optical_PCA.py:174-175 (Python banker's round)."""
    return max(MIN_SAMPLES_PCA, int(round(win_sec * fs_hz))), max(1, int(round(step_sec * fs_hz)))


def dynamic_pc1_sliding(time_sec, vx, vy, win_sec: float, step_sec: float, ref=np.array([0.0, 1.0]),
                        fs: float | None = None) -> np.ndarray:
    """Dynamic PC1 by sliding-window PCA on the GPU; same signature as optical_PCA.py:136-143.

    `fs` is explicit here (default: this module's global `fs`, as the reference reads its own global at
    optical_PCA.py:174-175) so a 60 fps clip does not need a module patch.  numpy in -> numpy out."""
    fs_hz = float(globals()["fs"] if fs is None else fs)
    time_sec = np.asarray(time_sec, dtype=float)
    n = int(time_sec.size)
    win_n, step_n = window_samples(win_sec, step_sec, fs_hz)
    out = pc1_sliding_batched(np.asarray(vx, float).reshape(1, -1), np.asarray(vy, float).reshape(1, -1),
                              [win_n], [step_n], ref)
    if out.shape[-1] != n:
        raise Cv2CompatError(-1, "time_sec, vx and vy must have the same length")
    return out[0, 0]


def pc1_sliding_batched(vx, vy, win_n: Sequence[int], step_n: Sequence[int], ref=(0.0, 1.0),
                        min_samples: int = MIN_SAMPLES_PCA):
    """PC1 for every (window configuration, series) pair in one set of launches.

    vx, vy: [n_series, n] float64 (numpy, or torch CUDA tensors to stay on the device).
    Returns [n_cfg, n_series, n]."""
    lib = _lib.load()
    ref = np.asarray(ref, float).reshape(2)
    win = np.ascontiguousarray(win_n, np.int32)
    step = np.ascontiguousarray(step_n, np.int32)
    if win.shape != step.shape or win.ndim != 1 or win.size < 1:
        raise Cv2CompatError(-1, "win_n and step_n must be equal-length 1-D sequences")
    is_torch = type(vx).__module__.startswith("torch")
    import torch
    if not torch.cuda.is_available():
        raise _lib.BtcsFlowError(_lib.BF_E_NODEVICE, "no CUDA device: dynamic_pc1_sliding has no CPU fallback")
    if is_torch:
        dvx = vx.to(torch.float64).contiguous()
        dvy = vy.to(torch.float64).contiguous()
    else:
        dvx = torch.from_numpy(np.ascontiguousarray(vx, np.float64)).cuda()
        dvy = torch.from_numpy(np.ascontiguousarray(vy, np.float64)).cuda()
    if dvx.dim() != 2 or dvx.shape != dvy.shape:
        raise Cv2CompatError(-1, "vx and vy must both be [n_series, n]")
    S, n = dvx.shape
    out = torch.full((win.size, S, n), float("nan"), dtype=torch.float64, device=dvx.device)
    if n > 0:
        stream = torch.cuda.current_stream(dvx.device).cuda_stream
        check(lib.bf_pc1_sliding_batched(dvx.data_ptr(), dvy.data_ptr(), S, n, win.ctypes.data, step.ctypes.data,
                                         win.size, float(ref[0]), float(ref[1]), int(min_samples), out.data_ptr(),
                                         stream))
    return out if is_torch else out.cpu().numpy()


# ---- host band-pass (optical_PCA.py:64-121) ---------------------------------------------------------------
def butter_bandpass_sos(low_hz: float, high_hz: float, fs: float, order: int = 4) -> np.ndarray:
    from scipy.signal import butter
    nyq = 0.5 * fs
    if not (0 < low_hz < high_hz < nyq):
        raise ValueError(f"Invalid band-pass range. low={low_hz}, high={high_hz}, nyquist={nyq}.")
    return butter(order, [low_hz / nyq, high_hz / nyq], btype="band", output="sos")


def sos_required_padlen(sos: np.ndarray) -> int:
    return 3 * (2 * int(sos.shape[0]))


def finite_runs(mask: np.ndarray) -> list[tuple[int, int]]:
    """Inclusive (start, end) of each run of True."""
    m = np.concatenate([[False], np.asarray(mask, bool), [False]])
    edges = np.flatnonzero(m[1:] != m[:-1])
    return [(int(a), int(b) - 1) for a, b in zip(edges[::2], edges[1::2])]


def bandpass_nanrobust(x: np.ndarray, sos: np.ndarray) -> np.ndarray:
    """Zero-phase band-pass on each finite run long enough for sosfiltfilt; the rest stays NaN."""
    from scipy.signal import sosfiltfilt
    x = np.asarray(x, dtype=float)
    y = np.full(x.shape, np.nan)
    need = sos_required_padlen(sos)
    for s, e in finite_runs(np.isfinite(x)):
        seg = x[s:e + 1]
        if seg.size <= need:
            continue
        pad = min(need, seg.size // 2 - 1)
        y[s:e + 1] = sosfiltfilt(sos, seg, padlen=pad) if pad > 0 else seg
    return y


def bandpass_nanrobust_device(x, sos: np.ndarray):
    """bandpass_nanrobust (optical_PCA.py:96-121) on the GPU: scipy.signal.sosfiltfilt per contiguous finite run,
    same length rules.  x: [n] or [n_series, n], numpy (returns numpy) or a torch CUDA tensor (stays on the device)."""
    import torch
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.BtcsFlowError(_lib.BF_E_NODEVICE, "no CUDA device: bandpass_nanrobust_device has no CPU fallback")
    sos = np.ascontiguousarray(sos, np.float64)
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise Cv2CompatError(-1, "sos must have shape (n_sections, 6)")
    is_torch = type(x).__module__.startswith("torch")
    dx = x.to(torch.float64) if is_torch else torch.from_numpy(np.ascontiguousarray(x, np.float64)).cuda()
    one_d = dx.dim() == 1
    dx = (dx[None] if one_d else dx).contiguous()
    out = torch.empty_like(dx)
    check(lib.bf_bandpass_nanrobust(dx.data_ptr(), dx.shape[0], dx.shape[1], sos.ctypes.data, None, sos.shape[0],
                                    out.data_ptr(), torch.cuda.current_stream(dx.device).cuda_stream))
    out = out[0] if one_d else out
    return out if is_torch else out.cpu().numpy()


def sosfilt_zi(sos: np.ndarray) -> np.ndarray:
    """scipy.signal.sosfilt_zi through the C library (host arithmetic; no GPU needed)."""
    sos = np.ascontiguousarray(sos, np.float64)
    zi = np.empty((sos.shape[0], 2), np.float64)
    check(_lib.load().bf_sosfilt_zi(sos.ctypes.data, sos.shape[0], zi.ctypes.data))
    return zi


def flow_to_pc1(t, vx, vy, fs_hz: float = fs, win_sec: float = WIN_SEC, step_sec: float = STEP_SEC,
                low_hz: float = BPF_LOW_HZ, high_hz: float = BPF_HIGH_HZ, order: int = BPF_ORDER,
                on_device: bool = True, sos: np.ndarray | None = None):
    """Band-pass + dynamic PC1 for one series: the body of optical_PCA.main (optical_PCA.py:254-267).

    on_device=True (default) runs the band-pass on the GPU too, so a torch CUDA series never leaves the device;
    on_device=False uses scipy on the host for the band-pass exactly like the reference.  `sos`: a filter already designed
    with butter_bandpass_sos (the design costs ~0.4 ms of host time; streaming callers design it once)."""
    if sos is None:
        sos = butter_bandpass_sos(low_hz, high_hz, fs_hz, order=order)
    if on_device:
        import torch
        is_torch = type(vx).__module__.startswith("torch")
        dvx = vx if is_torch else torch.from_numpy(np.ascontiguousarray(vx, np.float64)).cuda()
        dvy = vy if is_torch else torch.from_numpy(np.ascontiguousarray(vy, np.float64)).cuda()
        both = bandpass_nanrobust_device(torch.stack([dvx.to(torch.float64), dvy.to(torch.float64)]), sos)
        win_n, step_n = window_samples(win_sec, step_sec, fs_hz)
        out = pc1_sliding_batched(both[0:1], both[1:2], [win_n], [step_n], (0.0, 1.0))[0, 0]
        return out if is_torch else out.cpu().numpy()
    return dynamic_pc1_sliding(t, bandpass_nanrobust(vx, sos), bandpass_nanrobust(vy, sos), win_sec, step_sec,
                               np.array([0.0, 1.0]), fs=fs_hz)


def main(flow_csv: str = FLOW_CSV, out_csv: str = OUT_CSV) -> None:
    """flow.csv -> flow_pc1.csv with the reference's column names (optical_PCA.py:241-270)."""
    import pandas as pd
    df = pd.read_csv(flow_csv)
    required = {"t_sec", "vx_body", "vy_body"}
    missing = [c for c in required if c not in df.columns]
    if missing:
        raise KeyError(f"Missing columns in {flow_csv}. Required={sorted(required)}, missing={missing}.")
    t = df["t_sec"].to_numpy(float)
    pc1 = flow_to_pc1(t, df["vx_body"].to_numpy(float), df["vy_body"].to_numpy(float))
    pd.DataFrame({"t_sec": t, "pc1_dyn": pc1}).to_csv(out_csv, index=False)


if __name__ == "__main__":
    main()
