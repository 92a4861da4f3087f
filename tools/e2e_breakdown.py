"""Where the end-to-end (host buffers) step spends its time: PCIe rate, series call, PC1 tail."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import synthetic as syn, pca
spec, params = syn.config_spec("C2")
P = 256
spec.T = P + 1
dev = torch.device("cuda")
frames_dev = syn.make_clip(spec, dev, 0, P + 1)
frames_host = frames_dev.cpu().pin_memory().numpy()
mask_dev = torch.ones((1080, 1920), dtype=torch.uint8, device=dev)
mask_host = np.ones((1080, 1920), np.uint8)
plan = B.FlowPlan(1920, 1080, params, max_pairs=64)
def tm(fn, n=5):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t.append(1e3 * (time.perf_counter() - t0))
    return np.median(t)
stage = torch.empty_like(frames_dev)
hp = torch.from_numpy(frames_host)
ms = tm(lambda: stage.copy_(hp, non_blocking=True))
print(f"H2D of {frames_host.nbytes / 1e6:.0f} MB pinned: {ms:.2f} ms = {frames_host.nbytes / ms / 1e6:.1f} GB/s")
print(f"device-resident series call: {tm(lambda: plan.flow_series(frames_dev, None, None, mask_dev)):.2f} ms")
print(f"host-buffer series call:     {tm(lambda: plan.flow_series(frames_host, None, None, mask_host)):.2f} ms")
series = plan.flow_series(frames_host, None, None, mask_host)
def finish():
    s = torch.from_numpy(series).to(dev)[0].double()
    return pca.flow_to_pc1(None, s[:, 0].contiguous(), s[:, 1].contiguous(), fs_hz=30.0).cpu().numpy()
print(f"PC1 tail (series H2D, band-pass, PCA, PC1 D2H): {tm(finish):.2f} ms")
for first in (None,):
    pass
import os
for ramp in ("8,24", "64", "8", "16", "16,48", "4,12,36", "32"):
    os.environ["BTCSFLOW_HOST_CHUNKS"] = ramp
    print(f"host-buffer series call, chunk ramp {ramp:8s}: {tm(lambda: plan.flow_series(frames_host, None, None, mask_host), 7):.2f} ms")
os.environ.pop("BTCSFLOW_HOST_CHUNKS")
# how long do the pieces of the host path take when nothing overlaps?
t0 = time.perf_counter(); m = torch.from_numpy(mask_host).to(dev); torch.cuda.synchronize(); print(f"mask H2D: {1e3 * (time.perf_counter() - t0):.2f} ms")
