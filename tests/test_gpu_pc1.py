"""Sliding-window PCA -> PC1 on the GPU against the reference's outputs (golden) and the NumPy oracle, and the
end-to-end flow -> PC1 -> metrics parity gates of BASELINE.json (PC1 r >= 0.9999, same tau sign, ADS/AUC 0.1 %)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFGS = ((2.0, 0.1, 30), (0.5, 0.1, 30), (4.0, 0.25, 30), (1.0, 0.1, 60))


def test_pc1_matches_reference_golden(golden):
    from btcs_pnes_optical_flow_b200 import pca
    g = golden("pc1_golden.npz")
    t = g["t"]
    for s in range(3):
        vx, vy = g[f"vx{s}"], g[f"vy{s}"]
        for ws, ss, fs in CFGS:
            got = pca.dynamic_pc1_sliding(t, vx, vy, ws, ss, fs=fs)
            ref = g[f"pc1_{s}_{ws}_{ss}_{fs}"]
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            assert np.nanmax(np.abs(got - ref)) < 1e-12, (s, ws, np.nanmax(np.abs(got - ref)))
    assert np.isnan(pca.dynamic_pc1_sliding(t[:2], np.ones(2), np.ones(2), 2.0, 0.1)).all()
    assert np.isnan(pca.dynamic_pc1_sliding(t[:40], np.sin(t[:40]), np.cos(t[:40]), 2.0, 0.1)).all()
    nanv = np.full(100, np.nan)
    assert np.isnan(pca.dynamic_pc1_sliding(t[:100], nanv, nanv, 2.0, 0.1)).all()
    assert pca.dynamic_pc1_sliding(np.zeros(0), np.zeros(0), np.zeros(0), 2.0, 0.1).shape == (0,)


def test_pc1_random_series_vs_oracle():
    """Many seeded series with NaN gaps, long enough to need the multi-chunk scan (K > 1024 windows)."""
    from btcs_pnes_optical_flow_b200 import pca
    from oracle import pc1_np
    rng = np.random.default_rng(11)
    for trial in range(12):
        n = int(rng.integers(5, 5000))
        t = np.arange(n) / 30.0
        ang = rng.uniform(0, 3.1) + 0.8 * np.sin(0.3 * t)
        osc = np.sin(2 * np.pi * 2.5 * t + rng.uniform(0, 6))
        vx = osc * np.cos(ang) + 0.1 * rng.standard_normal(n)
        vy = osc * np.sin(ang) + 0.1 * rng.standard_normal(n)
        for _ in range(int(rng.integers(0, 6))):
            a = int(rng.integers(0, n))
            vx[a:a + int(rng.integers(1, 90))] = np.nan
        win_n, step_n = int(rng.integers(3, 130)), int(rng.integers(1, 7))
        ref = pc1_np.dynamic_pc1_sliding(vx, vy, win_n, step_n)
        got = pca.pc1_sliding_batched(vx[None], vy[None], [win_n], [step_n])[0, 0]
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (trial, n, win_n, step_n)
        if np.isfinite(ref).any():
            assert np.nanmax(np.abs(got - ref)) < 1e-11, (trial, n, win_n, step_n, np.nanmax(np.abs(got - ref)))


def test_pc1_batched_sweep_equals_single_calls(golden):
    """Config C5: several series x window sweep 0.5-4 s in one launch set."""
    import torch
    from btcs_pnes_optical_flow_b200 import pca
    g = golden("pc1_golden.npz")
    vx = np.stack([g[f"vx{s}"] for s in range(3)])
    vy = np.stack([g[f"vy{s}"] for s in range(3)])
    wins = [pca.window_samples(w, 0.1, 30) for w in (0.5, 1.0, 2.0, 4.0)]
    out = pca.pc1_sliding_batched(vx, vy, [w for w, _ in wins], [s for _, s in wins])
    assert out.shape == (4, 3, vx.shape[1])
    for c, (w, s) in enumerate(wins):
        for k in range(3):
            one = pca.pc1_sliding_batched(vx[k:k + 1], vy[k:k + 1], [w], [s])[0, 0]
            assert np.array_equal(out[c, k], one, equal_nan=True)
    dev = pca.pc1_sliding_batched(torch.from_numpy(vx).cuda(), torch.from_numpy(vy).cuda(), [60], [3])
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy()[0], out[2], equal_nan=True)


def test_end_to_end_flow_to_pc1_metrics_parity(golden):
    """Same clip as tests/golden/pipeline_golden.npz (seeded generator): GPU flow series -> host band-pass ->
    GPU PC1 -> host metrics, against the reference pipeline's golden outputs."""
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import metrics, pca, synthetic as syn
    g = golden("pipeline_golden.npz")
    spec = syn.ClipSpec(T=330, H=240, W=320, fps=30.0, seed=int(g["seed"]), patch=80, roi=100, amp=3.0, f0=3.0,
                        chirp=-0.05, tau=8.0)
    fr = syn.make_clip_np(spec)
    if not (np.array_equal(fr[:4], g["frames_head"]) and int(fr.astype(np.int64).sum()) == int(g["frames_sum"])):
        pytest.skip("synthetic generator is not bit-reproducible on this host; covered by the live-cv2 test below")
    rows = B.FlowPlan(320, 240, B.FB_PARAMS, max_pairs=16).flow_series(fr, None, None, g["mask"])[0]
    assert np.nanmax(np.abs(rows - g["rows"])) < 5e-4
    pc1 = pca.flow_to_pc1(g["t"], rows[:, 0].astype(float), rows[:, 1].astype(float))
    ok = np.isfinite(pc1) & np.isfinite(g["pc1"])
    assert np.array_equal(np.isfinite(pc1), np.isfinite(g["pc1"]))
    r = np.corrcoef(pc1[ok], g["pc1"][ok])[0, 1]
    assert r >= 0.9999, r
    m = metrics.compute_pc1_metrics(g["t"], pc1)
    assert np.sign(m["Kendall_tau_0_10"]) == np.sign(float(g["tau"]))
    assert abs(m["ADS_slope_0_10"] / float(g["ads"]) - 1) < 1e-3
    assert abs(m["PC1_area_0_10"] / float(g["area"]) - 1) < 1e-3


def test_end_to_end_against_live_cv2():
    """Config C1 geometry (640x480, 200x200 ROI), shortened to 12 s so the cv2 side takes seconds."""
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import metrics, pca, synthetic as syn
    from oracle import cv2_ref, pc1_np
    spec, p = syn.config_spec("C1", T=360)
    fr = syn.make_clip_np(spec)
    mask = spec.roi_mask()
    ref_rows = cv2_ref.roi_series(fr, [1.0, 0.0], [0.0, 1.0], mask, p)[0]
    rows = B.FlowPlan(spec.W, spec.H, p, max_pairs=16).flow_series(fr, None, None, mask)[0]
    assert np.nanmax(np.abs(rows - ref_rows)) < 5e-4
    t = np.arange(spec.T) / spec.fps
    sos = pca.butter_bandpass_sos(0.5, 5.0, 30)
    ref_pc1 = pc1_np.dynamic_pc1_sliding(pca.bandpass_nanrobust(ref_rows[:, 0], sos),
                                         pca.bandpass_nanrobust(ref_rows[:, 1], sos), 60, 3)
    pc1 = pca.flow_to_pc1(t, rows[:, 0].astype(float), rows[:, 1].astype(float))
    ok = np.isfinite(pc1) & np.isfinite(ref_pc1)
    assert ok.sum() > 300 and np.corrcoef(pc1[ok], ref_pc1[ok])[0, 1] >= 0.9999
    a, b = metrics.compute_pc1_metrics(t, pc1), metrics.compute_pc1_metrics(t, ref_pc1)
    assert np.sign(a["Kendall_tau_0_10"]) == np.sign(b["Kendall_tau_0_10"]) and b["Kendall_tau_0_10"] > 0.3
    assert abs(a["ADS_slope_0_10"] / b["ADS_slope_0_10"] - 1) < 1e-3
    assert abs(a["PC1_area_0_10"] / b["PC1_area_0_10"] - 1) < 1e-3


def test_device_bandpass_matches_scipy(golden):
    """bandpass_nanrobust on the GPU against the reference's outputs (golden) and the scipy host version."""
    import torch
    from btcs_pnes_optical_flow_b200 import pca
    g = golden("pc1_golden.npz")
    sos = pca.butter_bandpass_sos(0.5, 5.0, 30, order=4)
    x = np.stack([g[f"vx{s}"] for s in range(3)])
    got = pca.bandpass_nanrobust_device(x, sos)
    for s in range(3):
        ref = g[f"bp_vx{s}"]                                                  # output of the reference function itself
        assert np.array_equal(np.isnan(got[s]), np.isnan(ref))
        assert np.nanmax(np.abs(got[s] - ref)) < 1e-12, np.nanmax(np.abs(got[s] - ref))
    rng = np.random.default_rng(5)
    for n, order, band, fs in ((9000, 4, (0.5, 5.0), 30.0), (1234, 2, (1.0, 8.0), 60.0), (40, 4, (0.5, 5.0), 30.0),
                                (25, 4, (0.5, 5.0), 30.0), (24, 4, (0.5, 5.0), 30.0), (3000, 6, (0.3, 4.0), 25.0)):
        sos = pca.butter_bandpass_sos(band[0], band[1], fs, order=order)
        v = np.cumsum(rng.standard_normal(n)) * 0.1 + np.sin(np.arange(n) * 0.4)
        v[0] = np.nan
        for _ in range(4):
            a = int(rng.integers(0, n))
            v[a:a + int(rng.integers(1, 40))] = np.nan
        ref = pca.bandpass_nanrobust(v, sos)
        dev = pca.bandpass_nanrobust_device(torch.from_numpy(v).cuda(), sos)
        assert dev.is_cuda
        out = dev.cpu().numpy()
        assert np.array_equal(np.isnan(out), np.isnan(ref)), (n, order)
        if np.isfinite(ref).any():
            assert np.nanmax(np.abs(out - ref)) < 1e-11 * max(1.0, np.nanmax(np.abs(ref))), (n, order, np.nanmax(np.abs(out - ref)))
    # whole tail on the device == host band-pass + device PC1
    t = g["t"]
    a = pca.flow_to_pc1(t, g["vx0"], g["vy0"], on_device=True)
    b = pca.flow_to_pc1(t, g["vx0"], g["vy0"], on_device=False)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.nanmax(np.abs(a - b)) < 1e-11


def test_cli_pca_command_matches_reference_pipeline(tmp_path, golden):
    """`python -m btcs_pnes_optical_flow_b200 pca`: flow.csv -> flow_pc1.csv, against the reference's PC1 (golden)."""
    import pandas as pd
    from btcs_pnes_optical_flow_b200.__main__ import main
    g = golden("pipeline_golden.npz")
    rows = g["rows"]
    pd.DataFrame({"frame": np.arange(len(rows)), "t_sec": g["t"], "vx_body": rows[:, 0], "vy_body": rows[:, 1],
                  "mag_body": rows[:, 2]}).to_csv(tmp_path / "flow.csv", index=False)
    assert main(["pca", "--flow", str(tmp_path / "flow.csv"), "--out", str(tmp_path / "flow_pc1.csv")]) == 0
    df = pd.read_csv(tmp_path / "flow_pc1.csv")
    assert list(df.columns) == ["t_sec", "pc1_dyn"]
    got, ref = df["pc1_dyn"].to_numpy(), g["pc1"]
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.nanmax(np.abs(got - ref)) < 1e-10


def test_pc1_zero_dot_resets_the_sign_chain():
    """optical_PCA.py:203-205 flips a window's axis only when dot(w, prev_w) < 0.  When the dot is exactly 0 -- axis-aligned
    windows [1, 0] next to [0, 1], reachable when one component is constant in a window -- the reference keeps w as aligned
    to `ref`, whatever sign the chain carried.  (A prefix product of signs inherits the previous sign there.)"""
    from btcs_pnes_optical_flow_b200 import pca
    from oracle import pc1_np
    n, win_n, step_n = 200, 10, 10
    t = np.arange(n, dtype=float)
    osc = np.sin(0.9 * t)
    vx, vy = np.zeros(n), np.zeros(n)
    # windows (10 samples each): A) motion along (1, -1): aligned to ref=(0,1) it becomes (-1, 1)/sqrt2;
    # B) motion along x only -> sxy == 0 -> axis [1, 0], dot with A's signed axis < 0 -> flipped to [-1, 0] (chain sign -1);
    # C) motion along y only -> axis [0, 1]: dot with [-1, 0] is exactly 0 -> NOT flipped in the reference;
    # D) along x again, dot with [0, 1] exactly 0 -> stays [+1, 0]
    for k in range(0, n // win_n):
        s = slice(k * win_n, (k + 1) * win_n)
        kind = k % 4
        if kind == 0:
            vx[s], vy[s] = osc[s], -osc[s]
        elif kind == 1:
            vx[s] = osc[s]
        elif kind == 2:
            vy[s] = osc[s]
        else:
            vx[s] = osc[s]
    want = pc1_np.dynamic_pc1_sliding(vx, vy, win_n, step_n)
    got = pca.pc1_sliding_batched(vx[None], vy[None], [win_n], [step_n])[0, 0]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.nanmax(np.abs(got - want)) < 1e-12, np.nanmax(np.abs(got - want))
    assert np.abs(want).max() > 0.5
