"""Where does the largest endpoint error against cv2 sit?  (4K Gaussian config, compact and exact plans)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import btcs_pnes_optical_flow_b200 as B
from oracle import cv2_ref
from tests.helpers import textured
p = dict(pyr_scale=0.5, levels=5, winsize=21, iterations=3, poly_n=7, poly_sigma=1.5, flags=256)
a, b = textured(2160, 3840, 31), textured(2160, 3840, 31, shift=(3.4, -2.2))
ref = cv2_ref.farneback(a, b, **p)
for exact in (False, True):
    with B.FlowPlan(3840, 2160, p, max_pairs=1, exact=exact) as plan:
        got = plan.flow_pair(a, b)
    d = np.sqrt(((got - ref) ** 2).sum(-1))
    ys, xs = np.unravel_index(np.argsort(d, axis=None)[::-1][:8], d.shape)
    print("exact" if exact else "compact", "mean", d.mean(), "max", d.max())
    for y, x in zip(ys, xs):
        print(f"   ({x:4d},{y:4d}) err {d[y, x]:.4f}  got {got[y, x]}  ref {ref[y, x]}")
    print("   pixels with err > 5e-3:", int((d > 5e-3).sum()), " interior (16 px in):", int((d[16:-16, 16:-16] > 5e-3).sum()))
