"""Randomised parity sweep against cv2 on the GPU box: ragged sizes and parameter sets through flow_pair (compact and exact
plans) and through the series path.  Prints the worst cases; exits non-zero if a gate is broken.
usage: python tools/fuzz_parity.py [n_cases] [seed] [MAXWxMAXH]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from oracle import cv2_ref
from tests.helpers import textured, epe_banded

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
max_h, max_w = (int(v) for v in (sys.argv[3].split("x")[::-1] if len(sys.argv) > 3 else ("420", "520")))
worst = []
bad = 0
for case in range(n_cases):
    h = int(rng.integers(48, max_h)); w = int(rng.integers(48, max_w))
    if case % 3 == 0: w = (w // 4) * 4                      # the tile kernel needs w % 4 == 0; other widths take the generic path
    p = dict(B.FB_PARAMS)
    kind = case % 5
    if kind == 1: p.update(levels=int(rng.integers(1, 5)), iterations=int(rng.integers(1, 4)))
    if kind == 2: p.update(winsize=21, poly_n=7, poly_sigma=1.5, flags=256, levels=int(rng.integers(1, 5)))
    if kind == 3: p.update(pyr_scale=float(rng.choice([0.5, 0.6, 0.8])), winsize=int(rng.choice([9, 13, 15, 25])))
    if kind == 4: p.update(poly_n=int(rng.choice([5, 7])), poly_sigma=float(rng.choice([1.1, 1.2, 1.5])), flags=int(rng.choice([0, 256])))
    shift = (float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)))
    a, b = textured(h, w, 100 + case), textured(h, w, 100 + case, shift=shift)
    ref = cv2_ref.farneback(a, b, **p)
    band = 2 * (p["winsize"] // 2) + 2
    for exact in (False, True):
        with B.FlowPlan(w, h, p, max_pairs=2, exact=exact) as plan:
            got = plan.flow_pair(a, b)
            if not exact:
                inner_mask = np.zeros((h, w), np.uint8)
                if min(h, w) > 2 * band + 2: inner_mask[band:-band, band:-band] = 1
                else: inner_mask[:] = 1
                ser = plan.flow_series(torch.from_numpy(np.stack([a, b, a])).cuda(), None, None,
                                       torch.from_numpy(inner_mask).cuda()).cpu().numpy()[0]
        if min(h, w) <= 2 * band + 2:
            d = np.sqrt(((got - ref) ** 2).sum(-1)); mean, inner, edge = float(d.mean()), 0.0, float(d.max())
        else:
            mean, inner, edge = epe_banded(got, ref, band)
        # moving content over the whole frame (no static border): the north_star max gate is held on the WHOLE frame, border
        # band included, by both plans; interior far tighter
        ok = mean <= 1e-3 and inner <= 2e-3 and edge <= 0.05 and np.isfinite(got).all()
        if exact: ok = ok and inner <= 1e-4           # fp32 storage: ~1e-6 inside
        worst.append((inner, edge, mean, case, h, w, exact, p))
        if not ok:
            bad += 1
            print("GATE BROKEN", case, h, w, "exact" if exact else "compact", p, "mean %.2e inner %.2e band %.2e" % (mean, inner, edge))
    # series row 1 = ROI mean of the pair's flow (identity axes) over the interior (the border band may hold branch flips)
    want = ref[inner_mask != 0].mean(0, dtype=np.float64)   # float64: a float32 column-wise mean of 1e5 same-sign values drifts by ~1e-3
    if not (np.isnan(ser[0]).all() and np.abs(ser[1, :2] - want).max() < 5e-4):
        bad += 1
        print("SERIES MISMATCH", case, h, w, p, ser[1], want)
worst.sort(key=lambda t: -t[0])
print(f"{n_cases} cases x 2 plans; broken gates: {bad}")
for inner, edge, mean, case, h, w, exact, p in worst[:6]:
    print(f"  worst interior {inner:.2e} (band {edge:.2e}, mean {mean:.2e}) case {case} {w}x{h} {'exact' if exact else 'compact'} winsize {p['winsize']} poly_n {p['poly_n']} levels {p['levels']} flags {p['flags']} pyr_scale {p['pyr_scale']}")
sys.exit(1 if bad else 0)
