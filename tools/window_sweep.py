"""Every compiled window (winsize 4..33, box and Gaussian) on an odd-sized frame: tile kernels, compact and exact storage, against
the runtime-parameter fp32 kernel and cv2.  usage (GPU box): python tools/window_sweep.py"""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import btcs_pnes_optical_flow_b200 as B
from oracle import cv2_ref
from tests.helpers import textured, epe, epe_banded
h, w = 97, 131
a, b = textured(h, w, 5), textured(h, w, 5, shift=(0.9, -0.6))
for win in range(4, 34):
    for flags in (0, 256):
        p = dict(B.FB_PARAMS, winsize=win, levels=1, iterations=2, flags=flags)
        ref = cv2_ref.farneback(a, b, **p)
        os.environ["BTCSFLOW_NO_FAST"] = "1"
        with B.FlowPlan(w, h, p, exact=True) as plan: gen = plan.flow_pair(a, b)
        del os.environ["BTCSFLOW_NO_FAST"]
        out = []
        for exact in (False, True):
            with B.FlowPlan(w, h, p, exact=exact) as plan: got = plan.flow_pair(a, b)
            out.append("%s vs gen mean %.1e max %.1e | vs cv2 mean %.1e max %.1e" % ("exact" if exact else "compact", *epe(got, gen), *epe(got, ref)))
        print(win, flags, " || ".join(out), "|| gen vs cv2 max %.1e" % epe(gen, ref)[1], flush=True)
