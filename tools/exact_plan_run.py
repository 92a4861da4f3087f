"""One 32-pair 1080p series call on an exact (all-fp32 storage) plan, twice: the workload for ncu captures of the exact kernels."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import synthetic as syn
spec, params = syn.config_spec("C2")
spec.T = 33
dev = torch.device("cuda")
frames = syn.make_clip(spec, dev, 0, 33)
mask = torch.ones((1080, 1920), dtype=torch.uint8, device=dev)
plan = B.FlowPlan(1920, 1080, params, max_pairs=32, exact=True)
for _ in range(2):
    out = plan.flow_series(frames, None, None, mask)
torch.cuda.synchronize()
