"""CPU-side checks: host logic, and that the C-ABI library loads and exports every symbol include/btcsflow.h
declares.  No compute calls (there is no GPU here and no CPU fallback to call)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import _lib, flow, synthetic

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "btcsflow.h").read_text()
    declared = set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", header))
    declared -= {"bf_params", "bf_plan"}
    assert len(declared) >= 20
    lib = ctypes.CDLL(str(_lib.lib_path()))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in btcsflow.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert b"sm_100a" in _lib.load().bf_version()


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.BfParams) == 40   # double,int,int,int,int,double,int + tail padding
    assert _lib.BfParams.poly_sigma.offset == 24


def test_invalid_params_raise_like_cv2():
    a = np.zeros((64, 64), np.uint8)
    with pytest.raises(ValueError):            # cv2: -215 assertion (pyrScale_ < 1)
        B.calcOpticalFlowFarneback(a, a, None, **dict(B.FB_PARAMS, pyr_scale=1.0))
    with pytest.raises(ValueError):            # size mismatch
        B.calcOpticalFlowFarneback(a, a[:-1], None, **B.FB_PARAMS)
    with pytest.raises(ValueError):            # 3-channel input
        B.calcOpticalFlowFarneback(np.zeros((64, 64, 3), np.uint8), np.zeros((64, 64, 3), np.uint8), None, **B.FB_PARAMS)
    with pytest.raises(B.BtcsFlowError) as ei:   # valid for cv2, outside this library (validated before any device is touched)
        B.FlowPlan(64, 64, dict(B.FB_PARAMS, poly_n=17))
    assert ei.value.code == _lib.BF_E_UNSUPPORTED


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    a = np.zeros((64, 64), np.uint8)
    with pytest.raises(B.BtcsFlowError) as ei:
        B.calcOpticalFlowFarneback(a, a, None, **B.FB_PARAMS)
    assert ei.value.code == _lib.BF_E_NODEVICE
    with pytest.raises(B.BtcsFlowError):
        B.dynamic_pc1_sliding(np.arange(100) / 30.0, np.ones(100), np.ones(100), 2.0, 0.1)


def test_skel_index_and_roi_mask_match_reference(golden):
    g = golden("roi_golden.npz")
    got = [flow.skel_index_from_time(float(q), g["time_all"]) for q in g["sk_query"]]
    assert got == list(g["sk_index"])
    m = flow.build_roi_mask(120, 160, g["poly"])
    assert m.dtype == bool and np.array_equal(m, g["mask"])


def test_fb_params_are_the_reference_defaults():
    assert B.FB_PARAMS == dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    assert B.OPTFLOW_FARNEBACK_GAUSSIAN == 256 and B.OPTFLOW_USE_INITIAL_FLOW == 4


def test_synthetic_clip_is_deterministic_and_moves():
    spec = synthetic.ClipSpec(T=6, H=120, W=160, patch=48, roi=60, amp=3.0, seed=2)
    a = synthetic.make_clip_np(spec)
    b = synthetic.make_clip_np(spec)
    assert a.dtype == np.uint8 and a.shape == (6, 120, 160) and np.array_equal(a, b)
    assert np.array_equal(a[3:5], synthetic.make_clip_np(spec, 3, 2))     # chunked generation is consistent
    m = spec.roi_mask()
    assert m.sum() == 61 * 61
    assert (a[0] != a[2])[m].any() and not (a[0] != a[2])[~m].any()       # only the patch moves
    d = spec.displacement(np.arange(6) / 30.0)
    assert np.allclose(d[0], 0) and np.abs(d).max() <= 3.0
    for name in ("C1", "C2", "C3", "C4", "C5"):
        s, p = synthetic.config_spec(name)
        assert set(p) == set(B.FB_PARAMS)


def test_sosfilt_zi_matches_scipy():
    """bf_sosfilt_zi is host arithmetic inside the C library (no GPU): must equal scipy.signal.sosfilt_zi."""
    from scipy.signal import butter, sosfilt_zi
    from btcs_pnes_optical_flow_b200 import pca
    for order, band, fs in ((4, (0.5, 5.0), 30.0), (2, (1.0, 8.0), 60.0), (6, (0.3, 4.0), 25.0)):
        sos = butter(order, [band[0] / (fs / 2), band[1] / (fs / 2)], btype="band", output="sos")
        assert np.allclose(pca.sosfilt_zi(sos), sosfilt_zi(sos), rtol=1e-12, atol=1e-14)
    sos = butter(3, 0.2, output="sos")                                       # low-pass: non-zero DC gain chain
    assert np.allclose(pca.sosfilt_zi(sos), sosfilt_zi(sos), rtol=1e-12, atol=1e-14)


def test_bgr2gray_fixed_point_formula_matches_cv2():
    """Pins the 15-bit fixed-point formula the k_bgr2gray kernel implements against cv2.cvtColor (optical_flow.py:227)."""
    import cv2
    rng = np.random.default_rng(0)
    rand = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
    grid = np.stack(np.meshgrid(*[np.arange(0, 256, 5)] * 3, indexing="ij"), -1).reshape(-1, 1, 3).astype(np.uint8)
    for arr in (rand, grid):
        b, g, r = (arr[..., i].astype(np.int64) for i in range(3))
        mine = (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15
        assert np.array_equal(mine, cv2.cvtColor(arr, cv2.COLOR_BGR2GRAY))


def test_cli_parser_and_metrics_command(tmp_path, golden):
    """The CLI keeps the reference scripts' default file names; `metrics` runs on the host without a GPU."""
    import pandas as pd
    from btcs_pnes_optical_flow_b200.__main__ import build_parser, main, parse_roi
    ap = build_parser()
    a = ap.parse_args(["flow", "--video", "input.mp4", "--npz", "skeleton_pc1.npz", "--roi", "100,100;500,120;520,380;120,400"])
    assert a.out == "flow.csv" and a.roi.shape == (4, 2) and a.roi[2, 1] == 380.0
    assert ap.parse_args(["pca"]).flow == "flow.csv" and ap.parse_args(["pca"]).out == "flow_pc1.csv"
    assert ap.parse_args(["metrics"]).out == "flow_summary_dyn_core.csv"
    with pytest.raises(Exception):
        parse_roi("1,2;3,4")
    g = golden("pipeline_golden.npz")
    pd.DataFrame({"t_sec": g["t"], "pc1_dyn": g["pc1"]}).to_csv(tmp_path / "flow_pc1.csv", index=False)
    assert main(["metrics", "--pc1", str(tmp_path / "flow_pc1.csv"), "--out", str(tmp_path / "s.csv")]) == 0
    row = pd.read_csv(tmp_path / "s.csv").iloc[0]
    assert row["Peak_n"] == int(g["peak_n"]) and row["PC1_area_0_10"] == pytest.approx(float(g["area"]), rel=1e-9)
