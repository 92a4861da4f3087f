"""A/B timing of kernel variants selected by environment switches, one process, same frames.

usage: python tools/kernel_ab.py "" "BTCSFLOW_TILE_WARPS=6" "BTCSFLOW_TILE_TH=32,BTCSFLOW_X=1" ...
Prints device ms per 128-pair series call (median of NIT), the dominant kernel's average launch time and the max
difference of the ROI series against the first variant."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import synthetic as syn

variants = sys.argv[1:] or [""]
CFG = os.environ.get("KAB_CONFIG", "C2")          # KAB_CONFIG=C4 KAB_PAIRS=32 KAB_MAXPAIRS=16 for the 4K Gaussian case
spec, params = syn.config_spec(CFG)
P, NIT = int(os.environ.get("KAB_PAIRS", "128")), 7
MAXP = int(os.environ.get("KAB_MAXPAIRS", "64"))
EXACT = os.environ.get("KAB_EXACT", "0") == "1"
spec.T = P + 1
dev = torch.device("cuda")
frames = syn.make_clip(spec, dev, 0, P + 1)
mask = torch.ones((spec.H, spec.W), dtype=torch.uint8, device=dev)
base = None
touched = set()
for rep in range(2):                      # every variant twice, interleaved, to see drift
    for v in variants:
        for k in touched: os.environ.pop(k, None)
        for kv in filter(None, v.split(",")):
            k, val = kv.split("=")
            os.environ[k] = val
            touched.add(k)
        plan = B.FlowPlan(spec.W, spec.H, params, max_pairs=MAXP, exact=EXACT)
        plan.profile(True)
        ms = []
        for it in range(NIT):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            out = plan.flow_series(frames, None, None, mask)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2: ms.append(e0.elapsed_time(e1))
        pr = plan.profile_read()
        s = out[0].double().cpu().numpy()
        if base is None: base = s
        d = float(np.nanmax(np.abs(s - base)))
        st = {k: round(x["ms"] / max(x["launches"], 1), 3) for k, x in pr["stages"].items()}
        print(f"{v or 'default':40s} step {np.median(ms):7.2f} ms ({P / np.median(ms) * 1e3:6.0f} pairs/s)  avg launch ms {st}"
              f"  series diff vs first {d:.2e}", flush=True)
        plan.close()
