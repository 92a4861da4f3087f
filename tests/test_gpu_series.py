"""ROI body-axis series (the frame loop of optical_flow.py:218-250, batched on the GPU) against the
reference's own outputs (golden) and against cv2 on identical inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_series_matches_reference_golden(golden):
    """Golden rows come from /root/reference compute_roi_mean_body_flow with rotated axes and one NaN-axes frame."""
    import btcs_pnes_optical_flow_b200 as B
    g = golden("roi_golden.npz")
    fr, rows, ex, ey, mask = g["frames"], g["rows"], g["ex"], g["ey"], g["mask"]
    plan = B.FlowPlan(fr.shape[2], fr.shape[1], B.FB_PARAMS, max_pairs=4, max_rois=1)
    out = plan.flow_series(fr, ex, ey, mask)[0]
    assert out.shape == rows.shape
    assert np.isnan(out[0]).all() and np.isnan(out[4]).all()                      # row 0 and the NaN-axes row
    assert np.array_equal(np.isnan(out), np.isnan(rows))
    assert np.nanmax(np.abs(out - rows)) < 5e-4, np.nanmax(np.abs(out - rows))    # px; flow tolerance is 0.01
    # the per-pair reference function (optical_flow.py:136) gives the same three floats
    for t in (1, 2, 5):
        one = B.compute_roi_mean_body_flow(fr[t - 1], fr[t], ex[t], ey[t], mask, B.FB_PARAMS)
        assert np.allclose(one, out[t], rtol=0, atol=1e-6)
    plan.close()


def test_series_batching_multi_roi_and_device_api():
    import torch
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    from oracle import cv2_ref
    spec = syn.ClipSpec(T=11, H=135, W=240, seed=3, patch=60, roi=80, amp=3.0)
    fr = syn.make_clip_np(spec)
    m0 = spec.roi_mask()
    m1 = np.zeros_like(m0)
    m1[:, :120] = True
    empty = np.zeros_like(m0)
    masks = np.stack([m0, m1, empty])
    ref = cv2_ref.roi_series(fr, [1.0, 0.0], [0.0, 1.0], masks[:2], B.FB_PARAMS, threads=4)
    outs = []
    for mp in (1, 3, 16):                                                         # ring/batch boundaries
        plan = B.FlowPlan(240, 135, B.FB_PARAMS, max_pairs=mp, max_rois=3)
        out, flow = plan.flow_series(fr, None, None, masks, return_flow=True)
        outs.append(out)
        assert np.nanmax(np.abs(out[:2] - ref)) < 5e-4
        assert np.isnan(out[2]).all()                                             # empty ROI -> NaN like np.nanmean
        # dense flow of pair t equals the stand-alone pair call
        assert np.abs(flow[4] - B.calcOpticalFlowFarneback(fr[4], fr[5], None, **B.FB_PARAMS)).max() < 1e-5
        dev = plan.flow_series(torch.from_numpy(fr).cuda(), None, None, torch.from_numpy(masks).cuda())
        assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), out, equal_nan=True)   # host path == device path
        plan.close()
    assert np.array_equal(outs[0], outs[1], equal_nan=True) and np.array_equal(outs[0], outs[2], equal_nan=True)
    one = B.FlowPlan(240, 135, B.FB_PARAMS, max_pairs=2).flow_series(fr[:1])
    assert one.shape == (1, 1, 3) and np.isnan(one).all()                         # single frame: only the NaN row


def test_series_gaussian_poly7_config():
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    from oracle import cv2_ref
    p = dict(pyr_scale=0.5, levels=5, winsize=21, iterations=3, poly_n=7, poly_sigma=1.5, flags=256)   # config C4 params
    spec = syn.ClipSpec(T=5, H=272, W=480, seed=4, patch=90, roi=0, amp=4.0)
    fr = syn.make_clip_np(spec)
    ref = cv2_ref.roi_series(fr, [1.0, 0.0], [0.0, 1.0], spec.roi_mask(), p, threads=4)
    out = B.FlowPlan(480, 272, p, max_pairs=4).flow_series(fr)
    assert np.nanmax(np.abs(out - ref)) < 5e-4


def test_run_body_axis_flow_core_script_level(tmp_path):
    """Script-level drop-in: lossless FFV1 .avi + upstream NPZ -> flow.csv with the reference's columns."""
    import cv2
    import pandas as pd
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    from oracle import cv2_ref
    spec = syn.ClipSpec(T=20, H=120, W=160, seed=9, patch=48, roi=60, amp=3.0)
    fr = syn.make_clip_np(spec)
    vid = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(vid, cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (160, 120), isColor=True)
    if not wr.isOpened():
        pytest.skip("FFV1 writer unavailable")
    for f in fr:
        wr.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    wr.release()
    T = spec.T
    ex = np.tile([1.0, 0.0], (T, 1))
    ey = np.tile([0.0, 1.0], (T, 1))
    ex[7] = np.nan
    npz = str(tmp_path / "skel.npz")
    np.savez(npz, time_all=np.arange(T) / 30.0, fps=30.0, ex=ex, ey=ey)
    out_csv = str(tmp_path / "flow.csv")
    B.run_body_axis_flow_core(vid, npz, spec.roi_polygon(), out_csv, chunk_frames=6)                 # BGR->gray on the GPU
    host_csv = str(tmp_path / "flow_hostgray.csv")
    B.run_body_axis_flow_core(vid, npz, spec.roi_polygon(), host_csv, chunk_frames=7, gray_on_device=False)
    df = pd.read_csv(out_csv)
    assert np.allclose(df[["vx_body", "vy_body", "mag_body"]].to_numpy(),
                       pd.read_csv(host_csv)[["vx_body", "vy_body", "mag_body"]].to_numpy(), rtol=0, atol=1e-6, equal_nan=True)
    assert list(df.columns) == ["frame", "t_sec", "skel_idx", "axes_ok", "vx_body", "vy_body", "mag_body"]
    assert len(df) == T and df["vx_body"].isna().iloc[0]
    bad = df[df["axes_ok"] == 0]
    assert len(bad) >= 1 and bad["vx_body"].isna().all()
    # decode the same file on the host and run the cv2 reference over it
    cap = cv2.VideoCapture(vid)
    dec = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        dec.append(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY))
    dec = np.stack(dec)
    mask = B.build_roi_mask(120, 160, spec.roi_polygon())
    exr = ex[df["skel_idx"].to_numpy()]
    eyr = ey[df["skel_idx"].to_numpy()]
    ref = cv2_ref.roi_series(dec, exr, eyr, mask, B.FB_PARAMS, threads=4)[0]
    got = df[["vx_body", "vy_body", "mag_body"]].to_numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.nanmax(np.abs(got - ref)) < 5e-4


def test_bgr_to_gray_bit_exact():
    """cv2.cvtColor(BGR2GRAY) (optical_flow.py:227) on the GPU: bit-exact, odd widths and batches included."""
    import cv2
    import torch
    import btcs_pnes_optical_flow_b200 as B
    rng = np.random.default_rng(2)
    for shape in ((3, 120, 160, 3), (1, 67, 91, 3), (2, 1080, 1920, 3), (53, 77, 3)):
        x = rng.integers(0, 256, shape, dtype=np.uint8)
        got = B.bgr_to_gray(x)
        ref = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in x]) if x.ndim == 4 else cv2.cvtColor(x, cv2.COLOR_BGR2GRAY)
        assert got.dtype == np.uint8 and np.array_equal(got, ref)
    dev = B.bgr_to_gray(torch.from_numpy(x).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), ref)


def test_async_host_series_matches_sync():
    """flow_series_async: two clips in flight on one plan, results equal to the synchronous host call."""
    import torch
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    spec = syn.ClipSpec(T=21, H=135, W=240, seed=3, patch=60, roi=80, amp=3.0)
    a = syn.make_clip_np(spec)
    b = np.ascontiguousarray(a[::-1])
    pa, pb = torch.from_numpy(a).pin_memory().numpy(), torch.from_numpy(b).pin_memory().numpy()
    mask = spec.roi_mask()
    with B.FlowPlan(spec.W, spec.H, B.FB_PARAMS, max_pairs=4, max_rois=1) as plan:
        want_a, want_b = plan.flow_series(a, None, None, mask), plan.flow_series(b, None, None, mask)
        ha = plan.flow_series_async(pa, None, None, mask)
        hb = plan.flow_series_async(pb, None, None, mask)        # queued behind the first, staging buffers shared
        got_b, got_a = hb.result().copy(), ha.result().copy()
        one = plan.flow_series_async(pa[:1], None, None, mask).result()
    assert np.array_equal(got_a, want_a, equal_nan=True) and np.array_equal(got_b, want_b, equal_nan=True)
    assert one.shape == (1, 1, 3) and np.isnan(one).all()


def test_c2_full_size_series_batch_of_64_with_ring_wrap():
    """Config C2 as bench.py runs it: 1080p, 64 pairs per launch, a series longer than one batch (the frame ring wraps),
    full-frame ROI -- every row against cv2 on the same frames, and the dense field of pairs taken from inside a batch."""
    import torch
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    from oracle import cv2_ref
    spec, p = syn.config_spec("C2")
    T = 72
    spec.T = T
    fr = syn.make_clip(spec, "cuda", 0, T)
    fr_h = fr.cpu().numpy()
    check = [1, 2, 33, 63, 64, 65, 66, 71]                                     # both sides of the batch boundary (pair 64 = ring slot 0)
    with B.FlowPlan(spec.W, spec.H, p, max_pairs=64, max_rois=1) as plan:
        rows = plan.flow_series(fr, None, None, torch.from_numpy(spec.roi_mask()).cuda()).cpu().numpy()[0]
        _, flow = plan.flow_series(fr[:66], None, None, torch.from_numpy(spec.roi_mask()).cuda(), return_flow=True)
        dense = {k: flow[k].cpu().numpy() for k in (7, 40, 64)}
        del flow
    assert np.isnan(rows[0]).all() and np.isfinite(rows[1:]).all()
    mask = spec.roi_mask()
    for t in check:
        want = cv2_ref.roi_mean_body_flow(fr_h[t - 1], fr_h[t], [1.0, 0.0], [0.0, 1.0], mask, p)
        assert np.abs(rows[t] - np.asarray(want)).max() < 5e-4, (t, rows[t], want)
    from tests.helpers import epe_banded
    for k, got in dense.items():
        ref = cv2_ref.farneback(fr_h[k], fr_h[k + 1], **p)
        mean, inner, band = epe_banded(got, ref, 16)
        assert mean <= 1e-3 and inner <= 5e-3 and band <= 0.25, (k, mean, inner, band)   # static border: band bounded


def test_c5_two_rois_full_size():
    """Config C5: 1080p, two ROIs, rows of both against cv2; the window sweep runs in one batched PC1 call."""
    import torch
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import pca, synthetic as syn
    from oracle import cv2_ref, pc1_np
    spec, p = syn.config_spec("C5")
    T = 20
    spec.T = T
    fr = syn.make_clip(spec, "cuda", 0, T)
    fr_h = fr.cpu().numpy()
    masks = spec.roi_masks()
    assert masks.shape[0] == 2 and masks[0].sum() == masks[1].sum() > 0 and not (masks[0] & masks[1]).any()
    with B.FlowPlan(spec.W, spec.H, p, max_pairs=8, max_rois=2) as plan:
        rows = plan.flow_series(fr, None, None, torch.from_numpy(masks).cuda())
        ref = cv2_ref.roi_series(fr_h, [1.0, 0.0], [0.0, 1.0], masks, p, threads=8)
        got = rows.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.nanmax(np.abs(got - ref)) < 5e-4, np.nanmax(np.abs(got - ref))
        assert np.abs(ref[0, 1:, 2]).mean() > 10 * np.abs(ref[1, 1:, 2]).mean()          # ROI 0 holds the moving patch, ROI 1 background
    # the sweep: 2 ROIs x 4 windows in one call equals the sequential oracle per (ROI, window)
    n = 400
    t = np.arange(n) / 30.0
    vx = np.stack([np.sin(2 * np.pi * 3 * t) * 0.6, np.cos(2 * np.pi * 2 * t) * 0.3])
    vy = np.stack([np.sin(2 * np.pi * 3 * t + 0.2) * 0.8, np.sin(2 * np.pi * 2 * t) * 0.5])
    wins = [pc1_np.window_samples(w, 0.1, 30) for w in (0.5, 1.0, 2.0, 4.0)]
    out = pca.pc1_sliding_batched(vx, vy, [w for w, _ in wins], [s for _, s in wins])
    for c, (wn, sn) in enumerate(wins):
        for r in range(2):
            want = pc1_np.dynamic_pc1_sliding(vx[r], vy[r], wn, sn)
            assert np.array_equal(np.isnan(out[c, r]), np.isnan(want)) and np.nanmax(np.abs(out[c, r] - want)) < 1e-11
