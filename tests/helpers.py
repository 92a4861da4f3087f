"""Shared test helpers."""
import numpy as np


def textured(h, w, seed, shift=(0.0, 0.0)):
    """Smooth-noise texture with a sub-pixel translation (same recipe as tests/golden/make_golden.py)."""
    import cv2
    rng = np.random.default_rng(seed)
    big = rng.random((h + 64, w + 64)).astype(np.float32)
    big = cv2.GaussianBlur(big, (0, 0), 2.0)
    big = (big - big.min()) / (big.max() - big.min()) * 255
    M = np.array([[1, 0, -32 + shift[0]], [0, 1, -32 + shift[1]]], np.float32)
    out = cv2.warpAffine(big, M, (w, h), flags=cv2.INTER_CUBIC)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def epe(a, b):
    d = np.sqrt(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2).sum(-1))
    return float(d.mean()), float(d.max())


def epe_banded(a, b, band):
    """(mean over the frame, max over the interior `band` px in from every edge, max over the border band).

    The reference's UpdateMatrices switches branch where x + flow crosses the last row/column (`inside` test), so a 1e-4 px
    difference can flip a border pixel and the blur window spreads that over the band (true of any
    two builds of the algorithm, cv2 against the NumPy oracle included).  Parity is asserted tightly on the interior and
    on the mean, and bounded in the band."""
    d = np.sqrt(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2).sum(-1))
    inner = d[band:-band, band:-band]
    edge = d.copy()
    edge[band:-band, band:-band] = 0.0
    return float(d.mean()), float(inner.max()), float(edge.max())
