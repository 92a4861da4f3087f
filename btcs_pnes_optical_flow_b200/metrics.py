"""Host-side PC1 metrics: AUC of |PC1|, amplitude-decay slope (ADS), Kendall tau of inter-peak intervals.

These stay on the host by design (BASELINE.json north_star); they mirror /root/reference/optical_PC1.py
(helpers at :55-228, script body at :234-299).  The reference calls three functions it never defines
(`estimate_fs_from_time`, `safe_auc`, `exp_decay_regression`, optical_PC1.py:263/267/270 -- SURVEY Appendix D.1);
they are supplied here with the semantics its docstring and README describe, and parity tests apply the
*same* helpers to the cv2-derived and the GPU-derived series.
"""
from __future__ import annotations

import numpy as np

# Parameters with the reference's names and values (optical_PC1.py:33-44).
IN_CSV = "flow_pc1.csv"
OUT_CSV = "flow_summary_dyn_core.csv"
PC1_COL = "pc1_dyn"
WINDOW_SEC = 10.0
SMOOTH_SEC = 0.20
PEAK_MIN_FRAC = 0.20
PEAK_MIN_ABS = 0.0
MIN_DIST_SEC = 0.2


def ensure_odd(n: int) -> int:
    return int(n) | 1


# ---- the three helpers the reference forgot to ship -------------------------------------------------------
def estimate_fs_from_time(t: np.ndarray) -> float:
    """Sampling rate from the median time step."""
    t = np.asarray(t, float)
    dt = np.diff(t[np.isfinite(t)])
    dt = dt[dt > 0]
    if dt.size == 0:
        raise RuntimeError("cannot estimate fs: no positive time steps")
    return float(1.0 / np.median(dt))


def safe_auc(y: np.ndarray, t: np.ndarray) -> float:
    """Trapezoidal area over the samples where both y and t are finite."""
    y = np.asarray(y, float)
    t = np.asarray(t, float)
    m = np.isfinite(y) & np.isfinite(t)
    if int(m.sum()) < 2:
        return float("nan")
    return float(np.trapezoid(y[m], t[m]))


def exp_decay_regression(t: np.ndarray, amp: np.ndarray) -> dict:
    """Linear regression of ln(amp) on t over finite, positive samples -> slope (ADS), intercept, r, p, n."""
    from scipy.stats import linregress
    t = np.asarray(t, float)
    amp = np.asarray(amp, float)
    m = np.isfinite(t) & np.isfinite(amp) & (amp > 0)
    n = int(m.sum())
    if n < 3:
        nan = float("nan")
        return {"slope": nan, "intercept": nan, "r": nan, "p": nan, "n": n}
    res = linregress(t[m], np.log(amp[m]))
    return {"slope": float(res.slope), "intercept": float(res.intercept), "r": float(res.rvalue),
            "p": float(res.pvalue), "n": n}


# ---- smoothing / amplitude reference / cycle detection (optical_PC1.py:55-228) ------------------------------
def smooth_ma_nan(x: np.ndarray, fs: float, sec: float) -> np.ndarray:
    """Centred moving average of width ~sec that ignores NaNs (edge samples replicated)."""
    x = np.asarray(x, dtype=float)
    if sec <= 0:
        return x.copy()
    k = ensure_odd(max(1, int(round(fs * sec))))
    half = k // 2
    ok = np.isfinite(x)
    filled = np.where(ok, x, 0.0)
    box = np.ones(k) / k
    num = np.convolve(np.pad(filled, half, mode="edge"), box, mode="valid")
    den = np.convolve(np.pad(ok.astype(float), half, mode="edge"), box, mode="valid")
    y = num / np.maximum(den, 1e-12)
    y[den < 1e-12] = np.nan
    return y


def rolling_p95_positive(pc1_s: np.ndarray, fs: float, win_sec: float) -> np.ndarray:
    """Centred rolling 95th percentile of the positive part; NaN where fewer than 5 positive samples."""
    v = np.asarray(pc1_s, dtype=float)
    win_n = max(3, ensure_odd(int(round(win_sec * fs))))
    half = win_n // 2
    pos = np.where(np.isfinite(v) & (v > 0), v, np.nan)
    out = np.full(v.shape, np.nan)
    if v.size == 0:
        return out
    padded = np.pad(pos, half, mode="constant", constant_values=np.nan)
    win = np.lib.stride_tricks.sliding_window_view(padded, win_n)
    cnt = np.isfinite(win).sum(axis=1)
    good = cnt >= 5
    if good.any():
        with np.errstate(all="ignore"):
            out[good] = np.nanpercentile(win[good], 95, axis=1)
    return out


def detect_cycles_positive_peaks(pc1, time_sec, fs, smooth_sec: float = 0.20, p95_win_sec: float = 2.0,
                                 peak_min_frac: float = 0.20, peak_min_abs: float = 0.0, min_dist_sec: float = 0.2):
    """Positive-peak cycle detector: returns (pc1_s, t_peaks, tm, T) like optical_PC1.py:121-228.

    Cycle = upward zero crossing .. next downward zero crossing of the smoothed waveform; its maximum is
    the peak; a peak survives if it reaches max(peak_min_abs, peak_min_frac * local p95); peaks closer
    than min_dist_sec are merged keeping the larger."""
    pc1 = np.asarray(pc1, dtype=float)
    time_sec = np.asarray(time_sec, dtype=float)
    pc1_s = smooth_ma_nan(pc1, fs, smooth_sec)
    ref = rolling_p95_positive(pc1_s, fs, win_sec=p95_win_sec)
    a, b = pc1_s[:-1], pc1_s[1:]
    with np.errstate(invalid="ignore"):
        ups = np.flatnonzero((a <= 0) & (b > 0))
        downs = np.flatnonzero((a > 0) & (b <= 0))
    peaks_t: list[float] = []
    peaks_a: list[float] = []
    for iu in ups:
        k = np.searchsorted(downs, iu, side="right")
        if k >= downs.size:
            continue
        end = int(downs[k])
        seg = pc1_s[iu:end + 1]
        if seg.size == 0 or not np.isfinite(seg).any():
            continue
        im = int(np.nanargmax(seg))
        amp = float(seg[im])
        ipk = int(iu) + im
        thr = float(peak_min_abs)
        r = ref[ipk]
        if np.isfinite(r) and r > 0:
            thr = max(thr, float(peak_min_frac) * float(r))
        if amp < thr:
            continue
        peaks_t.append(float(time_sec[ipk]))
        peaks_a.append(amp)
    empty = np.array([])
    if len(peaks_t) < 2:
        return pc1_s, np.asarray(peaks_t, float), empty, empty
    kept_t = [peaks_t[0]]
    kept_a = [peaks_a[0]]
    for t, amp in zip(peaks_t[1:], peaks_a[1:]):
        if t - kept_t[-1] < float(min_dist_sec):
            if amp > kept_a[-1]:
                kept_t[-1], kept_a[-1] = t, amp
        else:
            kept_t.append(t)
            kept_a.append(amp)
    t_peaks = np.asarray(kept_t, float)
    if t_peaks.size < 2:
        return pc1_s, t_peaks, empty, empty
    T = np.diff(t_peaks)
    tm = 0.5 * (t_peaks[:-1] + t_peaks[1:])
    ok = np.isfinite(T) & (T > 0)
    return pc1_s, t_peaks, tm[ok], T[ok]


def compute_pc1_metrics(t_sec: np.ndarray, pc1: np.ndarray, window_sec: float = WINDOW_SEC) -> dict:
    """The script body of optical_PC1.py:241-299 as a function: one summary row as a dict."""
    from scipy.stats import kendalltau
    t_all = np.asarray(t_sec, float)
    p_all = np.asarray(pc1, float)
    m = np.isfinite(t_all) & np.isfinite(p_all)
    t_all, p_all = t_all[m], p_all[m]
    if t_all.size < 10:
        raise RuntimeError("Too few valid samples in input CSV.")
    time = t_all - float(t_all[0])
    w = (time >= 0.0) & (time <= float(window_sec))
    time, p = time[w], p_all[w]
    if time.size < 10:
        raise RuntimeError("Too few samples in the 0–10 s window.")
    fs_est = estimate_fs_from_time(time)
    amp = smooth_ma_nan(np.abs(p), fs_est, SMOOTH_SEC)
    area = safe_auc(amp, time)
    ads = exp_decay_regression(time, amp)
    r2 = float(ads["r"] ** 2) if np.isfinite(ads["r"]) else float("nan")
    _, t_peaks, tm, T = detect_cycles_positive_peaks(p, time, fs_est, smooth_sec=SMOOTH_SEC,
                                                     peak_min_frac=PEAK_MIN_FRAC, peak_min_abs=PEAK_MIN_ABS,
                                                     min_dist_sec=MIN_DIST_SEC)
    if tm.size >= 5:
        tau, pval = kendalltau(tm, T)
        tau, pval = float(tau), float(pval)
    else:
        tau = pval = float("nan")
    return {
        "PC1_source": PC1_COL,
        "window_sec": float(window_sec),
        "PC1_area_0_10": float(area),
        "ADS_slope_0_10": float(ads["slope"]),
        "ADS_R2_0_10": r2,
        "Kendall_tau_0_10": tau,
        "Kendall_p_0_10": pval,
        "Peak_n": int(t_peaks.size),
    }


def main(in_csv: str = IN_CSV, out_csv: str = OUT_CSV) -> None:
    """flow_pc1.csv -> one-row summary CSV with the reference's column names (optical_PC1.py:285-299)."""
    import pandas as pd
    df = pd.read_csv(in_csv)
    required = {"t_sec", PC1_COL}
    missing = [c for c in required if c not in df.columns]
    if missing:
        raise KeyError(f"Missing columns in {in_csv}. Required={sorted(required)}, missing={missing}.")
    row = compute_pc1_metrics(df["t_sec"].to_numpy(float), df[PC1_COL].to_numpy(float))
    pd.DataFrame([row]).to_csv(out_csv, index=False)


if __name__ == "__main__":
    main()
