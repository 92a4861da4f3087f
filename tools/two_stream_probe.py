"""Does running two independent plans on two streams (alternate 64-pair chunks of the same clip) beat one plan?  Full-grid
kernels rarely overlap usefully; this measures it.  usage (GPU box): python tools/two_stream_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import synthetic as syn

spec, params = syn.config_spec("C2")
P = 256
spec.T = P + 1
dev = torch.device("cuda")
frames = syn.make_clip(spec, dev, 0, P + 1)
mask = torch.ones((1080, 1920), dtype=torch.uint8, device=dev)
half = P // 2
fa, fb = frames[: half + 1].contiguous(), frames[half:].contiguous()


def timed(fn, nit=6):
    ms = []
    for it in range(nit):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 2: ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


one = B.FlowPlan(1920, 1080, params, max_pairs=64)
t1 = timed(lambda: one.flow_series(frames, None, None, mask))
print(f"one plan, one stream : {t1:7.2f} ms  {P / t1 * 1e3:6.0f} pairs/s")
two = [B.FlowPlan(1920, 1080, params, max_pairs=64) for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]


def both():
    cur = torch.cuda.current_stream()
    for s in streams: s.wait_stream(cur)
    for pl, s, f in zip(two, streams, (fa, fb)):
        with torch.cuda.stream(s):
            pl.flow_series(f, None, None, mask)
    for s in streams: cur.wait_stream(s)


t2 = timed(both)
print(f"two plans, two streams: {t2:7.2f} ms  {P / t2 * 1e3:6.0f} pairs/s")
