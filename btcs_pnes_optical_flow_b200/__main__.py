"""Command line around the three stages, with the reference scripts' default file names.

    python -m btcs_pnes_optical_flow_b200 flow    --video input.mp4 --npz skeleton_pc1.npz \\
                                                  --roi "100,100;500,120;520,380;120,400" [--out flow.csv]
    python -m btcs_pnes_optical_flow_b200 pca     [--flow flow.csv] [--out flow_pc1.csv]
    python -m btcs_pnes_optical_flow_b200 metrics [--pc1 flow_pc1.csv] [--out flow_summary_dyn_core.csv]

flow    = optical_flow.py  (run_body_axis_flow_core, /root/reference/optical_flow.py:195-259; example values :265-288)
pca     = optical_PCA.py   (main, optical_PCA.py:241-270)
metrics = optical_PC1.py   (script body, optical_PC1.py:234-299)
"""
from __future__ import annotations

import argparse

import numpy as np


def parse_roi(text: str) -> np.ndarray:
    """ "x,y;x,y;..." -> float array [N, 2] (any N >= 3; the reference docstring says 4, its code takes any N)."""
    pts = [tuple(float(v) for v in p.split(",")) for p in text.split(";") if p.strip()]
    arr = np.asarray(pts, dtype=float)
    if arr.ndim != 2 or arr.shape[1] != 2 or arr.shape[0] < 3:
        raise argparse.ArgumentTypeError("ROI must be 'x,y;x,y;x,y[;...]' with at least 3 vertices")
    return arr


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="python -m btcs_pnes_optical_flow_b200", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    f = sub.add_parser("flow", help="video + body axes -> flow.csv (GPU Farneback + ROI means)")
    f.add_argument("--video", required=True)
    f.add_argument("--npz", required=True, help="upstream NPZ with time_all, fps, ex, ey")
    f.add_argument("--roi", required=True, type=parse_roi, help="polygon 'x,y;x,y;...'")
    f.add_argument("--out", default="flow.csv")
    f.add_argument("--chunk-frames", type=int, default=64)
    p = sub.add_parser("pca", help="flow.csv -> flow_pc1.csv (band-pass + sliding-window PCA on the GPU)")
    p.add_argument("--flow", default="flow.csv")
    p.add_argument("--out", default="flow_pc1.csv")
    m = sub.add_parser("metrics", help="flow_pc1.csv -> one-row summary (AUC, ADS, Kendall tau; host)")
    m.add_argument("--pc1", default="flow_pc1.csv")
    m.add_argument("--out", default="flow_summary_dyn_core.csv")
    return ap


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.cmd == "flow":
        from .flow import run_body_axis_flow_core
        run_body_axis_flow_core(args.video, args.npz, args.roi, args.out, chunk_frames=args.chunk_frames)
    elif args.cmd == "pca":
        from . import pca
        pca.main(args.flow, args.out)
    else:
        from . import metrics
        metrics.main(args.pc1, args.out)
    print("Saved:", args.out)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
