"""Where does the streaming host-buffer path lose time against the device-resident path?  (a) device-resident calls back to
back, (b) host-buffer async calls back to back, all queued then waited, (c) the bench's pipelined loop (launch i+1, then the
PC1 tail of i), (d) the same without the PC1 tail."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import synthetic as syn, pca
spec, params = syn.config_spec("C2")
P, K = 256, 10
spec.T = P + 1
dev = torch.device("cuda")
frames_dev = syn.make_clip(spec, dev, 0, P + 1)
frames_host = frames_dev.cpu().pin_memory().numpy()
mask_dev = torch.ones((1, 1080, 1920), dtype=torch.uint8, device=dev)
mask_host = np.ones((1, 1080, 1920), np.uint8)
plan = B.FlowPlan(1920, 1080, params, max_pairs=64)
sos = pca.butter_bandpass_sos(0.5, 5.0, 30.0)
side = torch.cuda.Stream()

def tail(series):
    with torch.cuda.stream(side):
        s = torch.from_numpy(series).to(dev).double()
        both = pca.bandpass_nanrobust_device(torch.cat([s[:, :, 0], s[:, :, 1]]), sos)
        return pca.pc1_sliding_batched(both[:1].contiguous(), both[1:].contiguous(), [60], [3]).cpu().numpy()

def timeit(fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3

def a():
    for _ in range(K): plan.flow_series(frames_dev, None, None, mask_dev)
def b():
    hs = [plan.flow_series_async(frames_host, None, None, mask_host) for _ in range(K)]
    for h in hs: h.result()
def c(with_tail=True):
    pend = None
    for _ in range(K):
        nxt = plan.flow_series_async(frames_host, None, None, mask_host)
        if pend is not None:
            r = pend.result()
            if with_tail: tail(r)
        pend = nxt
    r = pend.result()
    if with_tail: tail(r)
for rep in range(2):
    print(f"device-resident back to back {timeit(a):7.2f} ms/step | host async all queued {timeit(b):7.2f} | pipelined + PC1 tail {timeit(c):7.2f} | pipelined, no tail {timeit(lambda: c(False)):7.2f}", flush=True)
t0 = time.perf_counter(); h = plan.flow_series_async(frames_host, None, None, mask_host); t1 = time.perf_counter(); h.result()
print(f"host time inside one flow_series_async call: {(t1 - t0) * 1e3:.2f} ms")
