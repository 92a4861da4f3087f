"""Per-step wall/device time of the device-resident series call, to look for host-side stalls."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import synthetic as syn, pca
spec, params = syn.config_spec("C2")
P = 128
spec.T = P + 1
dev = torch.device("cuda")
frames = syn.make_clip(spec, dev, 0, P + 1)
mask = torch.ones((1080, 1920), dtype=torch.uint8, device=dev)
plan = B.FlowPlan(1920, 1080, params, max_pairs=16)
plan.profile(True)
NIT = int(sys.argv[1]) if len(sys.argv) > 1 else 14
for it in range(NIT):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = plan.flow_series(frames, None, None, mask)
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    s = out[0].double().cpu().numpy()
    t3 = time.perf_counter()
    print(f"step {it:2d}: launch {1e3*(t1-t0):7.2f} ms  total {1e3*(t2-t0):7.2f} ms  device {e0.elapsed_time(e1):7.2f} ms  d2h {1e3*(t3-t2):5.2f} ms")

pr = plan.profile_read(); print("dominant kernel avg ms", pr["total_ms"] / max(pr["launches"], 1), "launches", pr["launches"])
if NIT < 14: sys.exit(0)
print("---- PC1 stage (host band-pass + GPU PC1) on a 2049-sample series")
n = 2049
t = np.arange(n) / 30.0
s = np.sin(2 * np.pi * 3 * t)[:, None] * np.array([0.6, 0.8])[None] + 0.01 * np.random.default_rng(0).standard_normal((n, 2))
s[0] = np.nan
sos = pca.butter_bandpass_sos(0.5, 5.0, 30.0)
for it in range(12):
    t0 = time.perf_counter()
    a, b = pca.bandpass_nanrobust(s[:, 0], sos), pca.bandpass_nanrobust(s[:, 1], sos)
    t1 = time.perf_counter()
    r = pca.dynamic_pc1_sliding(t, a, b, 2.0, 0.1, fs=30.0)
    t2 = time.perf_counter()
    print(f"iter {it:2d}: band-pass {1e3*(t1-t0):7.2f} ms   pc1 {1e3*(t2-t1):7.2f} ms")
