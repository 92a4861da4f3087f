"""Device-resident flow-series throughput for one BASELINE config (C1..C5): pairs/s, flow stage only."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 16
mp = int(sys.argv[3]) if len(sys.argv) > 3 else 8
spec, params = syn.config_spec(name)
spec.T = P + 1
dev = torch.device("cuda")
frames = syn.make_clip(spec, dev, 0, P + 1)
nroi = 2 if name == "C5" else 1
mask = torch.ones((nroi, spec.H, spec.W), dtype=torch.uint8, device=dev)
plan = B.FlowPlan(spec.W, spec.H, params, max_pairs=mp, max_rois=nroi)
print(name, params, "scales", [(s["w"], s["h"]) for s in plan.scales()], "coeff bits", plan.coeff_storage_bits, "workspace MiB", plan.workspace_bytes >> 20)
for it in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = plan.flow_series(frames, None, None, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"  iter {it}: {ms:8.2f} ms for {P} pairs -> {P / ms * 1e3:8.1f} pairs/s")
