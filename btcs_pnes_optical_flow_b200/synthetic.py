"""Synthetic rhythmic-motion clips (BASELINE.json configs).

A static smooth-noise background plus a textured square patch that translates along a fixed direction with
displacement A * exp(-t / tau) * sin(phi(t)), phi(t) = 2 pi (f0 t + chirp t^2): a decaying, slowing
oscillation (clonic-like).  The slowing chirp makes inter-peak intervals grow, so Kendall's tau is clearly
positive instead of a coin flip.  Written with torch ops only, so the same code generates on
the CPU (tests) or directly in HBM (bench).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class ClipSpec:
    T: int = 900
    H: int = 480
    W: int = 640
    fps: float = 30.0
    seed: int = 0
    patch: int = 160          # side of the moving textured square (px)
    roi: int = 200            # side of the square ROI centred on the patch (0 = full frame)
    amp: float = 6.0          # peak displacement (px)
    f0: float = 3.0           # Hz at t = 0
    chirp: float = -0.05      # Hz/s (negative = slowing)
    tau: float = 8.0          # amplitude decay constant (s)
    direction: tuple = (0.6, 0.8)   # unit vector (x, y) of the motion
    center: tuple | None = None     # patch centre (x, y); default frame centre
    roi2_dx: int = 0                # != 0: a second ROI of the same size, its centre shifted by this many pixels in x

    def displacement(self, t: np.ndarray) -> np.ndarray:
        t = np.asarray(t, float)
        d = self.amp * np.exp(-t / self.tau) * np.sin(2 * np.pi * (self.f0 * t + self.chirp * t * t))
        return d[:, None] * np.asarray(self.direction, float)[None, :]

    def roi_polygon(self) -> np.ndarray:
        cx, cy = self.center or (self.W // 2, self.H // 2)
        if self.roi <= 0:
            return np.array([[0, 0], [self.W - 1, 0], [self.W - 1, self.H - 1], [0, self.H - 1]], float)
        r = self.roi // 2
        return np.array([[cx - r, cy - r], [cx + r, cy - r], [cx + r, cy + r], [cx - r, cy + r]], float)

    def roi_mask(self) -> np.ndarray:
        m = np.zeros((self.H, self.W), bool)
        if self.roi <= 0:
            m[:] = True
            return m
        cx, cy = self.center or (self.W // 2, self.H // 2)
        r = self.roi // 2
        m[max(cy - r, 0):cy + r + 1, max(cx - r, 0):cx + r + 1] = True  # inclusive like cv2.fillPoly
        return m

    def roi_masks(self) -> np.ndarray:
        """[n_roi, H, W] bool: the ROI around the patch and, for two-ROI configurations (C5: bilateral limbs), a second ROI
        of the same size beside it."""
        masks = [self.roi_mask()]
        if self.roi2_dx and self.roi > 0:
            cx, cy = self.center or (self.W // 2, self.H // 2)
            masks.append(ClipSpec(H=self.H, W=self.W, roi=self.roi, center=(cx + self.roi2_dx, cy)).roi_mask())
        return np.stack(masks)


def smooth_noise(h: int, w: int, gen: torch.Generator, device, lo: float = 30.0, hi: float = 225.0) -> torch.Tensor:
    """Band-limited texture in [lo, hi]: three octaves of bicubically upsampled uniform noise."""
    acc = torch.zeros((1, 1, h, w), device=device)
    for cell, wgt in ((24, 0.45), (9, 0.35), (4, 0.20)):
        gh, gw = max(h // cell, 2) + 2, max(w // cell, 2) + 2
        g = torch.rand((1, 1, gh, gw), generator=gen, device=device)
        acc += wgt * F.interpolate(g, size=(h, w), mode="bicubic", align_corners=True)
    acc = (acc - acc.min()) / (acc.max() - acc.min()).clamp_min(1e-6)
    return (lo + (hi - lo) * acc)[0, 0]


def make_clip(spec: ClipSpec, device="cpu", t_start: int = 0, t_count: int | None = None) -> torch.Tensor:
    """uint8 frames [t_count, H, W] for frame indices t_start .. t_start + t_count - 1 of the clip."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(spec.seed))
    H, W, P = spec.H, spec.W, spec.patch
    bg = smooth_noise(H, W, gen, dev)
    margin = int(math.ceil(spec.amp)) + 3
    tex = smooth_noise(P + 2 * margin, P + 2 * margin, gen, dev, lo=10.0, hi=245.0)
    cx, cy = spec.center or (W // 2, H // 2)
    x0, y0 = cx - P // 2 - margin, cy - P // 2 - margin        # window that can contain the patch
    x1, y1 = x0 + P + 2 * margin, y0 + P + 2 * margin
    if x0 < 0 or y0 < 0 or x1 > W or y1 > H:
        raise ValueError("patch (plus motion margin) does not fit in the frame")
    n = spec.T - t_start if t_count is None else t_count
    tt = (np.arange(t_start, t_start + n) / spec.fps)
    disp = torch.as_tensor(spec.displacement(tt), dtype=torch.float32, device=dev)  # [n, 2] (dx, dy)
    frames = bg.round().clamp(0, 255).to(torch.uint8).unsqueeze(0).repeat(n, 1, 1)
    # window pixel centres relative to the undisplaced patch origin
    ys = torch.arange(y0, y1, device=dev, dtype=torch.float32) - (cy - P // 2)
    xs = torch.arange(x0, x1, device=dev, dtype=torch.float32) - (cx - P // 2)
    bgw = bg[y0:y1, x0:x1]
    size = P + 2 * margin
    chunk = max(1, min(n, (64 << 20) // (size * size * 16)))
    for s in range(0, n, chunk):
        d = disp[s:s + chunk]
        m = d.shape[0]
        px = xs[None, None, :] - d[:, 0, None, None]            # patch-local coordinates of each pixel
        py = ys[None, :, None] - d[:, 1, None, None]
        px = px.expand(m, size, size)
        py = py.expand(m, size, size)
        inside = (px >= 0) & (px <= P - 1) & (py >= 0) & (py <= P - 1)
        gx = (px + margin) / (size - 1) * 2 - 1
        gy = (py + margin) / (size - 1) * 2 - 1
        grid = torch.stack([gx, gy], dim=-1)
        samp = F.grid_sample(tex[None, None].expand(m, 1, size, size), grid, mode="bilinear",
                             padding_mode="border", align_corners=True)[:, 0]
        win = torch.where(inside, samp, bgw[None])
        frames[s:s + m, y0:y1, x0:x1] = win.round().clamp(0, 255).to(torch.uint8)
    return frames


def make_clip_np(spec: ClipSpec, t_start: int = 0, t_count: int | None = None) -> np.ndarray:
    return make_clip(spec, "cpu", t_start, t_count).numpy()


# The five BASELINE.json configurations.
def config_spec(name: str, **over) -> tuple[ClipSpec, dict]:
    """(clip spec, Farneback params) for 'C1' .. 'C5'."""
    base = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    if name == "C1":
        spec = ClipSpec(T=900, H=480, W=640, fps=30.0, patch=160, roi=200, amp=6.0)
    elif name in ("C2", "C3", "C5"):
        spec = ClipSpec(T=9000 if name == "C2" else 300, H=1080, W=1920, fps=30.0, patch=360, roi=0 if name != "C5" else 440,
                        amp=8.0, center=(700, 540) if name == "C5" else None, roi2_dx=520 if name == "C5" else 0)
    elif name == "C4":
        spec = ClipSpec(T=600, H=2160, W=3840, fps=60.0, patch=720, roi=0, amp=12.0)
        base.update(levels=5, winsize=21, poly_n=7, poly_sigma=1.5, flags=256)
    else:
        raise KeyError(name)
    for k, v in over.items():
        setattr(spec, k, v)
    return spec, base
