"""Kernel-by-kernel parity of the CUDA stages (through the C-ABI) against the stage-level oracle."""
import numpy as np
import pytest

from tests.helpers import textured

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def _dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape,levels,pyr", [((135, 240), 3, 0.5), ((200, 264), 4, 0.7), ((97, 131), 2, 0.5)])
def test_level_images(T, shape, levels, pyr):
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import stages
    from oracle import farneback_np as fb
    h, w = shape
    img = textured(h, w, 3)
    plan = B.FlowPlan(w, h, dict(B.FB_PARAMS, levels=levels, pyr_scale=pyr))
    scales = fb.select_scales(w, h, pyr, levels)
    got_sc = plan.scales()
    assert [(s.w, s.h, s.ksize) for s in scales] == [(s["w"], s["h"], s["ksize"]) for s in got_sc]
    for i, sc in enumerate(scales):
        ref = fb.level_image(img, sc)
        got = stages.level_image(plan, _dev(T, img), i).cpu().numpy()
        assert np.abs(got - ref).max() < 1e-3, (i, np.abs(got - ref).max())          # on 0..255 data
        got_f = stages.level_image(plan, _dev(T, img.astype(np.float32)), i).cpu().numpy()
        assert np.array_equal(got, got_f)                                            # f32 input == u8 input
    plan.close()


@pytest.mark.parametrize("n,sigma,w", [(5, 1.2, 150), (5, 1.2, 152), (7, 1.5, 152), (7, 1.5, 131), (3, 1.1, 152), (9, 0.0, 150)])
def test_poly_exp(T, n, sigma, w):
    """w % 4 == 0 takes the compile-time-N kernel (n = 5, 7), anything else the runtime-parameter one."""
    from btcs_pnes_optical_flow_b200 import stages
    from oracle import farneback_np as fb
    img = textured(101, w, 5).astype(np.float32)
    ref = fb.poly_exp(img, n, sigma)
    got = stages.poly_exp(_dev(T, img), n, sigma).cpu().numpy().transpose(1, 2, 0)
    scale = np.abs(ref).max(axis=(0, 1))
    assert (np.abs(got - ref).max(axis=(0, 1)) < 2e-5 * scale + 1e-5).all(), np.abs(got - ref).max(axis=(0, 1))


def test_update_matrices(T):
    from btcs_pnes_optical_flow_b200 import stages
    from oracle import farneback_np as fb
    rng = np.random.default_rng(0)
    h, w = 70, 93
    a, b = textured(h, w, 6).astype(np.float32), textured(h, w, 6, shift=(1.3, 0.6)).astype(np.float32)
    R0, R1 = fb.poly_exp(a, 5, 1.2), fb.poly_exp(b, 5, 1.2)
    flow = (rng.standard_normal((h, w, 2)) * 3).astype(np.float32)
    flow[0, :, 0] = -5.0        # pushes the footprint outside -> fallback branch
    flow[:, -1, 0] = 0.0        # x1 = w-1 -> outside by the unsigned test
    flow[10, 10] = (1e9, -1e9)  # absurd flow must land in the fallback branch, not crash
    ref = fb.update_matrices(R0, R1, flow)
    got = stages.update_matrices(_dev(T, R0.transpose(2, 0, 1)), _dev(T, R1.transpose(2, 0, 1)), _dev(T, flow))
    got = got.cpu().numpy().transpose(1, 2, 0)
    m = np.ones((h, w), bool)
    m[10, 10] = False           # float32 (x + 1e9) products differ in rounding order; branch is what matters
    tol = 1e-5 * np.abs(ref[m]).max(axis=0) + 1e-6
    assert (np.abs(got - ref)[m].max(axis=0) < tol).all(), np.abs(got - ref)[m].max(axis=0)
    assert np.isfinite(got).all()


@pytest.mark.parametrize("winsize,flags,w", [(15, 0, 120), (15, 0, 301), (15, 0, 300), (16, 0, 120), (5, 0, 120),
                                              (21, 256, 120), (16, 256, 120), (33, 0, 120)])
def test_blur_solve(T, winsize, flags, w):
    """winsize 15 box with w % 4 == 0 takes the specialised kernel (tiles of 128 x 32: 300 and 120 exercise the
    ragged right/bottom edges); everything else the runtime-parameter kernel."""
    from btcs_pnes_optical_flow_b200 import stages
    from oracle import farneback_np as fb
    h = 83
    a, b = textured(h, w, 8).astype(np.float32), textured(h, w, 8, shift=(0.7, -1.1)).astype(np.float32)
    R0, R1 = fb.poly_exp(a, 5, 1.2), fb.poly_exp(b, 5, 1.2)
    M = fb.update_matrices(R0, R1, np.zeros((h, w, 2), np.float32))
    blur = fb.blur_gauss if flags else fb.blur_box
    ref = fb.solve_flow(blur(M, winsize))
    got = stages.blur_solve(_dev(T, M.transpose(2, 0, 1)), winsize, flags).cpu().numpy()
    assert np.abs(got - ref).max() < 2e-4, np.abs(got - ref).max()


def test_upsample_flow(T):
    from btcs_pnes_optical_flow_b200 import stages
    from oracle import farneback_np as fb
    rng = np.random.default_rng(1)
    f = rng.standard_normal((34, 60, 2)).astype(np.float32)
    for (w, h) in ((120, 68), (85, 49), (60, 34)):
        ref = fb.resize_bilinear(f, w, h) * np.float32(2.0)
        got = stages.upsample_flow(_dev(T, f), w, h, 2.0).cpu().numpy()
        assert np.abs(got - ref).max() < 1e-5


def test_no_out_of_bounds_writes(T):
    """Every stage kernel on awkward sizes (ragged tiles, odd widths) with canary bands around the outputs."""
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import stages
    rng = np.random.default_rng(3)
    stages.check_guards()
    for (h, w) in ((24, 128), (25, 132), (83, 120), (97, 131), (50, 300), (33, 36)):
        img = rng.random((h, w)).astype(np.float32) * 255
        for n in (5, 7, 3):
            stages.poly_exp(_dev(T, img), n, 1.2)
        R = rng.standard_normal((5, h, w)).astype(np.float32)
        fl = (rng.standard_normal((h, w, 2)) * 4).astype(np.float32)
        M = stages.update_matrices(_dev(T, R), _dev(T, R[::-1].copy()), _dev(T, fl))
        for ws, flags in ((15, 0), (21, 256), (9, 0), (16, 256)):
            stages.blur_solve(M, ws, flags)
        stages.upsample_flow(_dev(T, fl), 2 * w + 1, 2 * h - 1, 2.0)
        if min(h, w) >= 32:
            plan = B.FlowPlan(w, h, dict(B.FB_PARAMS, levels=2))
            for i in range(len(plan.scales())):
                stages.level_image(plan, _dev(T, img.astype(np.uint8)), i)
            plan.close()
    assert stages.check_guards() > 50
