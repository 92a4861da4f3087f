"""CPU study of storage formats for the update matrices M and the polynomial coefficients R: the NumPy oracle
(oracle/farneback_np.py) run with the values rounded the way a storage format would round them, compared with cv2.
This is where the compact plan's layout comes from:

    all16   M = (G11, G12, G22, h1, h2) all fp16 (round 1):           ~1e-4 px mean, up to 3e-3 interior
    g16h32  G fp16, h fp32 (independent rounding):                     no better -- rounding G alone is as harmful
    cons    G fp16, h = A b + Gq d formed from the ROUNDED G, fp32:    ~4e-7 .. 2e-6 mean, <= 2e-4 interior (exact-grade)
    qR      R's quadratic terms fp16 (+ fp16 bilinear weights):        ~1e-6 mean

usage: python tools/emulate_storage.py [n_cases]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from oracle import cv2_ref, farneback_np as F
from tests.helpers import epe_banded, textured

f32 = np.float32


def q16(x):
    return x.astype(np.float16).astype(np.float32)


def update(R0, R1, flow, mode, w16=False):
    """UpdateMatrices (oracle/farneback_np.update_matrices) with the storage rounding of `mode` applied to the result."""
    h, w = flow.shape[:2]
    X, Y = np.meshgrid(np.arange(w), np.arange(h))
    dx, dy = flow[..., 0].astype(f32), flow[..., 1].astype(f32)
    fx, fy = X.astype(f32) + dx, Y.astype(f32) + dy
    x1, y1 = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
    ax, ay = (fx - x1.astype(f32)).astype(f32), (fy - y1.astype(f32)).astype(f32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xs, ys = np.clip(x1, 0, max(w - 2, 0)), np.clip(y1, 0, max(h - 2, 0))
    one = f32(1)
    wts = [(one - ax) * (one - ay), ax * (one - ay), (one - ax) * ay, ax * ay]
    taps = [R1[ys, xs], R1[ys, xs + 1], R1[ys + 1, xs], R1[ys + 1, xs + 1]]
    Rw = np.zeros((h, w, 5), f32)
    for a, t in zip(wts, taps):
        Rw[..., :2] += a[..., None] * t[..., :2]
        Rw[..., 2:] += (q16(a) if w16 else a)[..., None] * t[..., 2:]
    r4 = np.where(inside, (R0[..., 2] + Rw[..., 2]) * f32(0.5), R0[..., 2])
    r5 = np.where(inside, (R0[..., 3] + Rw[..., 3]) * f32(0.5), R0[..., 3])
    r6 = np.where(inside, (R0[..., 4] + Rw[..., 4]) * f32(0.25), R0[..., 4] * f32(0.5))
    b2 = (R0[..., 0] - np.where(inside, Rw[..., 0], 0)) * f32(0.5)
    b3 = (R0[..., 1] - np.where(inside, Rw[..., 1], 0)) * f32(0.5)
    sx, sy = np.ones(w, f32), np.ones(h, f32)
    for i in range(5):
        sx[i] *= F._BORDER[i]; sx[w - 1 - i] *= F._BORDER[i]; sy[i] *= F._BORDER[i]; sy[h - 1 - i] *= F._BORDER[i]
    sc = (sx[None, :] * sy[:, None]).astype(f32)
    b2, b3, r4, r5, r6 = (v * sc for v in (b2, b3, r4, r5, r6))
    G = np.stack([r4 * r4 + r6 * r6, (r4 + r5) * r6, r5 * r5 + r6 * r6], -1)
    Gh = q16(G) if mode in ("cons", "cons16") else G            # the G that h is formed from
    h1 = r4 * b2 + r6 * b3 + Gh[..., 0] * dy + Gh[..., 1] * dx
    h2 = r6 * b2 + r5 * b3 + Gh[..., 1] * dy + Gh[..., 2] * dx
    if mode in ("all16", "g16h32", "cons", "cons16"):
        G = q16(G)
    if mode in ("all16", "cons16"):
        h1, h2 = q16(h1), q16(h2)
    return np.concatenate([G, h1[..., None], h2[..., None]], -1).astype(f32)


def farneback(prev, nxt, mode="exact", qR=False, w16=False, **p):
    H, W = prev.shape
    blur = F.blur_gauss if (p["flags"] & 256) else F.blur_box
    cur = None
    for sc in F.select_scales(W, H, p["pyr_scale"], p["levels"]):
        cur = np.zeros((sc.h, sc.w, 2), f32) if cur is None else F.resize_bilinear(cur, sc.w, sc.h) * f32(1.0 / p["pyr_scale"])
        R0 = F.poly_exp(F.level_image(prev, sc), p["poly_n"], p["poly_sigma"])
        R1 = F.poly_exp(F.level_image(nxt, sc), p["poly_n"], p["poly_sigma"])
        if qR:
            R0[..., 2:], R1[..., 2:] = q16(R0[..., 2:]), q16(R1[..., 2:])
        M = update(R0, R1, cur, mode, w16)
        for i in range(p["iterations"]):
            cur = F.solve_flow(blur(M, p["winsize"]))
            if i < p["iterations"] - 1:
                M = update(R0, R1, cur, mode, w16)
    return cur


if __name__ == "__main__":
    rng = np.random.default_rng(9)
    for case in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
        h, w = int(rng.integers(48, 300)), int(rng.integers(48, 700))
        p = dict(cv2_ref.FB_PARAMS)
        if case % 3 == 1:
            p.update(winsize=15, flags=256)
        if case % 3 == 2:
            p.update(winsize=21, poly_n=7, poly_sigma=1.5, flags=256, levels=int(rng.integers(1, 5)))
        shift = (float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)))
        a, b = textured(h, w, 100 + case), textured(h, w, 100 + case, shift=shift)
        ref = cv2_ref.farneback(a, b, **p)
        band = 2 * (p["winsize"] // 2) + 2
        cells = []
        for name, kw in (("exact", {}), ("all16", dict(mode="all16")), ("g16h32", dict(mode="g16h32")), ("cons", dict(mode="cons")),
                         ("cons16", dict(mode="cons16")), ("qR+w16", dict(qR=True, w16=True)),
                         ("compact", dict(mode="cons", qR=True, w16=True))):
            cells.append("%s %.1e/%.1e/%.1e" % (name, *epe_banded(farneback(a, b, **kw, **p), ref, band)))
        print(f"{w}x{h} winsize {p['winsize']} flags {p['flags']} | mean/interior max/band max: " + " | ".join(cells), flush=True)
