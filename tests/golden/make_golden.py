"""Generate the committed golden fixtures from the REFERENCE ITSELF (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/optical_flow.py and optical_PCA.py by path (both have __main__ guards), runs
optical_PC1.py unmodified through runpy with the three helpers it forgets to define injected
(SURVEY Appendix D.1), and calls the installed cv2 (the un-vendored dependency that holds the hot-path
arithmetic).  /root/reference does not exist on the GPU box, so tests only ever read the .npz files
written here.
"""
from __future__ import annotations

import importlib.util
import os
import runpy
import sys
import tempfile
from pathlib import Path

import cv2
import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from btcs_pnes_optical_flow_b200 import metrics as my_metrics  # noqa: E402  (the three restored helpers)
from btcs_pnes_optical_flow_b200 import synthetic as syn  # noqa: E402


def load_ref(name: str):
    spec = importlib.util.spec_from_file_location(f"ref_{name}", REF / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def textured(h, w, seed, shift=(0.0, 0.0)):
    rng = np.random.default_rng(seed)
    big = rng.random((h + 64, w + 64)).astype(np.float32)
    big = cv2.GaussianBlur(big, (0, 0), 2.0)
    big = (big - big.min()) / (big.max() - big.min()) * 255
    M = np.array([[1, 0, -32 + shift[0]], [0, 1, -32 + shift[1]]], np.float32)
    out = cv2.warpAffine(big, M, (w, h), flags=cv2.INTER_CUBIC)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


FLOW_CASES = {
    "defaults": dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0),
    "gauss7": dict(pyr_scale=0.5, levels=5, winsize=21, iterations=3, poly_n=7, poly_sigma=1.5, flags=256),
    "single": dict(pyr_scale=0.5, levels=0, winsize=15, iterations=1, poly_n=5, poly_sigma=1.2, flags=0),
    "odd": dict(pyr_scale=0.7, levels=4, winsize=16, iterations=2, poly_n=3, poly_sigma=1.1, flags=0),
}


def farneback_golden():
    a = textured(96, 128, 11)
    b = textured(96, 128, 11, shift=(1.6, -0.9))
    c = textured(67, 91, 12)
    d = textured(67, 91, 12, shift=(-2.3, 1.2))
    out = dict(a=a, b=b, c=c, d=d, cv2_version=np.array(cv2.__version__))
    for name, p in FLOW_CASES.items():
        out[f"flow_ab_{name}"] = cv2.calcOpticalFlowFarneback(a, b, None, **p)
        out[f"flow_cd_{name}"] = cv2.calcOpticalFlowFarneback(c, d, None, **p)
    np.savez_compressed(HERE / "farneback_golden.npz", **out)


def roi_golden(of):
    spec = syn.ClipSpec(T=7, H=120, W=160, fps=30.0, seed=5, patch=48, roi=60, amp=3.0, f0=3.0)
    frames = syn.make_clip_np(spec)
    poly = np.array([[40.7, 30.2], [120.9, 36.5], [118.1, 95.8], [44.3, 90.0]])
    mask = of.build_roi_mask(spec.H, spec.W, poly)
    th = 0.3
    ex = np.tile(np.array([np.cos(th), np.sin(th)]), (spec.T, 1))
    ey = np.tile(np.array([-np.sin(th), np.cos(th)]), (spec.T, 1))
    ex[4] = np.nan                                   # invalid axes -> NaN row, prev still advances
    rows = np.full((spec.T, 3), np.nan)
    for t in range(1, spec.T):
        if np.isfinite(ex[t]).all() and np.isfinite(ey[t]).all():
            rows[t] = of.compute_roi_mean_body_flow(frames[t - 1], frames[t], ex[t], ey[t], mask, of.FB_PARAMS)
    time_all = np.array([0.0, 0.1, 0.25, 0.4])
    q = np.array([-1.0, 0.0, 0.05, 0.1, 0.24, 0.25, 0.39, 0.4, 9.0])
    sk = np.array([of.skel_index_from_time(float(v), time_all) for v in q])
    np.savez_compressed(HERE / "roi_golden.npz", frames=frames, poly=poly, mask=mask, ex=ex, ey=ey, rows=rows,
                        time_all=time_all, sk_query=q, sk_index=sk)


def pc1_golden(pca):
    out = {}
    rng = np.random.default_rng(3)
    n = 420
    t = np.arange(n) / 30.0
    for s in range(3):
        ph = rng.uniform(0, 6.28)
        ang = 0.9 + 0.6 * np.sin(0.4 * t + s)        # slowly rotating principal direction
        osc = np.exp(-t / 9) * np.sin(2 * np.pi * (3 * t - 0.05 * t * t) + ph)
        vx = osc * np.cos(ang) + 0.05 * rng.standard_normal(n)
        vy = osc * np.sin(ang) + 0.05 * rng.standard_normal(n)
        vx[0] = vy[0] = np.nan
        for _ in range(3):
            a = int(rng.integers(5, n - 40))
            L = int(rng.integers(1, 30))
            vx[a:a + L] = np.nan
            if rng.random() < 0.5:
                vy[a + 2:a + L + 4] = np.nan
        out[f"vx{s}"], out[f"vy{s}"] = vx, vy
        for (ws, ss, fs) in ((2.0, 0.1, 30), (0.5, 0.1, 30), (4.0, 0.25, 30), (1.0, 0.1, 60)):
            pca.fs = fs
            out[f"pc1_{s}_{ws}_{ss}_{fs}"] = pca.dynamic_pc1_sliding(t, vx, vy, ws, ss)
        pca.fs = 30
        sos = pca.butter_bandpass_sos(0.5, 5.0, 30, order=4)
        out[f"bp_vx{s}"] = pca.bandpass_nanrobust(vx, sos)
    out["t"] = t
    # degenerate inputs
    out["pc1_short"] = pca.dynamic_pc1_sliding(t[:2], np.ones(2), np.ones(2), 2.0, 0.1)
    out["pc1_lt_win"] = pca.dynamic_pc1_sliding(t[:40], np.sin(t[:40]), np.cos(t[:40]), 2.0, 0.1)
    nanv = np.full(100, np.nan)
    out["pc1_allnan"] = pca.dynamic_pc1_sliding(t[:100], nanv, nanv, 2.0, 0.1)
    np.savez_compressed(HERE / "pc1_golden.npz", **out)


def pipeline_golden(of, pca):
    """C1-like clip at quarter size: reference flow series -> reference band-pass + PC1 -> reference metrics."""
    spec = syn.ClipSpec(T=330, H=240, W=320, fps=30.0, seed=7, patch=80, roi=100, amp=3.0, f0=3.0, chirp=-0.05,
                        tau=8.0)
    frames = syn.make_clip_np(spec)
    mask = of.build_roi_mask(spec.H, spec.W, spec.roi_polygon())
    ex, ey = np.array([1.0, 0.0]), np.array([0.0, 1.0])
    rows = np.full((spec.T, 3), np.nan)
    for t in range(1, spec.T):
        rows[t] = of.compute_roi_mean_body_flow(frames[t - 1], frames[t], ex, ey, mask, of.FB_PARAMS)
    tt = np.arange(spec.T) / spec.fps
    pca.fs = 30
    sos = pca.butter_bandpass_sos(pca.BPF_LOW_HZ, pca.BPF_HIGH_HZ, pca.fs, order=pca.BPF_ORDER)
    vx_b = pca.bandpass_nanrobust(rows[:, 0], sos)
    vy_b = pca.bandpass_nanrobust(rows[:, 1], sos)
    pc1 = pca.dynamic_pc1_sliding(tt, vx_b, vy_b, pca.WIN_SEC, pca.STEP_SEC, np.array([0.0, 1.0]))
    with tempfile.TemporaryDirectory() as td:
        cwd = os.getcwd()
        os.chdir(td)
        try:
            pd.DataFrame({"t_sec": tt, "pc1_dyn": pc1}).to_csv("flow_pc1.csv", index=False)
            runpy.run_path(str(REF / "optical_PC1.py"), init_globals={
                "estimate_fs_from_time": my_metrics.estimate_fs_from_time,
                "safe_auc": my_metrics.safe_auc,
                "exp_decay_regression": my_metrics.exp_decay_regression,
            })
            summ = pd.read_csv("flow_summary_dyn_core.csv").iloc[0]
        finally:
            os.chdir(cwd)
    np.savez_compressed(
        HERE / "pipeline_golden.npz", seed=spec.seed, rows=rows, t=tt, vx_b=vx_b, vy_b=vy_b, pc1=pc1, mask=mask,
        frames_head=frames[:4], frames_sum=np.array(int(frames.astype(np.int64).sum())),
        area=float(summ["PC1_area_0_10"]), ads=float(summ["ADS_slope_0_10"]), ads_r2=float(summ["ADS_R2_0_10"]),
        tau=float(summ["Kendall_tau_0_10"]), tau_p=float(summ["Kendall_p_0_10"]), peak_n=int(summ["Peak_n"]))
    print("pipeline summary:", dict(summ))


if __name__ == "__main__":
    of = load_ref("optical_flow")
    pca = load_ref("optical_PCA")
    farneback_golden()
    roi_golden(of)
    pc1_golden(pca)
    pipeline_golden(of, pca)
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size)
