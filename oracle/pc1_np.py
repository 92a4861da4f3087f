"""CPU restatement (NumPy, float64) of the reference's sliding-window PCA -> PC1 -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/optical_PCA.py:136-235 (`dynamic_pc1_sliding`) step by step, with `fs` as an explicit
argument instead of the module global the reference reads at optical_PCA.py:174-175.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.

Pinning: ``tests/golden/pc1_golden.npz`` holds outputs of the *reference function itself* (imported from
/root/reference in the build container by ``tests/golden/make_golden.py``) on seeded series with NaN gaps;
``tests/test_oracle_pc1.py`` checks this restatement against them (<= 1e-12).
"""
from __future__ import annotations

import numpy as np

MIN_SAMPLES_PCA = 3  # optical_PCA.py:58


def window_samples(win_sec: float, step_sec: float, fs: float) -> tuple[int, int]:
    """optical_PCA.py:174-175 (Python round = half to even)."""
    return max(MIN_SAMPLES_PCA, int(round(win_sec * fs))), max(1, int(round(step_sec * fs)))


def principal_axis(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Unit eigenvector of the largest eigenvalue of cov([x, y]) (optical_PCA.py:192-199)."""
    X = np.column_stack([x, y])
    Xc = X - X.mean(axis=0)
    C = (Xc.T @ Xc) / max(len(x) - 1, 1)
    vals, vecs = np.linalg.eigh(C)
    return vecs[:, int(np.argmax(vals))]


def dynamic_pc1_sliding(vx, vy, win_n: int, step_n: int, ref=(0.0, 1.0), min_samples: int = MIN_SAMPLES_PCA):
    """pc1_dyn[n] for window length / step given in samples."""
    vx = np.asarray(vx, float)
    vy = np.asarray(vy, float)
    ref = np.asarray(ref, float)
    n = vx.size
    out = np.full(n, np.nan)
    if n < min_samples:
        return out
    centres, axes = [], []
    prev = None
    start = 0
    while start + win_n <= n:                                    # optical_PCA.py:181
        sx, sy = vx[start:start + win_n], vy[start:start + win_n]
        ok = np.isfinite(sx) & np.isfinite(sy)                   # :187
        if ok.sum() >= min_samples:                              # :188
            w = principal_axis(sx[ok], sy[ok])
            if np.dot(w, ref) < 0:                               # :127-133, :202
                w = -w
            if prev is not None and np.dot(w, prev) < 0:         # :203-204
                w = -w
            prev = w
            centres.append((2 * start + win_n - 1) // 2)         # :207
            axes.append(w)
        start += step_n
    if not centres:
        return out
    c = np.asarray(centres)
    A = np.vstack(axes)
    i = np.arange(n)
    j = np.clip(np.searchsorted(c, i, side="left"), 0, len(c) - 1)   # :218-219
    j2 = np.maximum(j - 1, 0)
    pick = np.where(np.abs(i - c[j2]) < np.abs(i - c[j]), j2, j)     # :221-225 (ties -> later centre)
    ok = np.isfinite(vx) & np.isfinite(vy)
    out[ok] = vx[ok] * A[pick[ok], 0] + vy[ok] * A[pick[ok], 1]      # :227-233
    return out
