"""Phase timeline of k_blur_solve_box: rebuilds the library with -DBF_TRACE (debug stamps at the phase boundaries), runs
one 64-pair finest-scale workload and reports phase durations and how co-resident CTAs overlap.  Run on the GPU box:
    python tools/trace_phases.py [ENV=VAL ...]        -> gpurun_out/trace_phases.txt (+ .npy)"""
import ctypes as C, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    os.environ[k] = v
from btcs_pnes_optical_flow_b200 import build as b
b.NVCC_FLAGS.append("-DBF_TRACE")
b.build(force=True)
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import _lib, synthetic as syn

lib = _lib.load()
lib.bf_debug_trace_set.restype = C.c_int
lib.bf_debug_trace_set.argtypes = [C.c_void_p]
spec, params = syn.config_spec("C2")
P = 64
spec.T = P + 1
dev = torch.device("cuda")
frames = syn.make_clip(spec, dev, 0, P + 1)
mask = torch.ones((1080, 1920), dtype=torch.uint8, device=dev)
plan = B.FlowPlan(1920, 1080, params, max_pairs=P)
for _ in range(3):
    plan.flow_series(frames, None, None, mask)
torch.cuda.synchronize()
ncta = 15 * 45 * P
buf = torch.zeros(ncta * 8, dtype=torch.int64, device=dev)
assert lib.bf_debug_trace_set(buf.data_ptr()) == 0
plan.flow_series(frames, None, None, mask)       # every level stamps; the finest scale's last launches overwrite (same grid)
torch.cuda.synchronize()
lib.bf_debug_trace_set(None)
t = buf.cpu().numpy().reshape(ncta, 8)
out = Path(ROOT / "gpurun_out"); out.mkdir(exist_ok=True)
np.save(out / "trace_phases.npy", t)
# only launches with the update tail stamp; the survivors are the second finest-scale iteration
d = np.diff(t[:, :5].astype(np.int64), axis=1)
lines = []
lines.append(f"CTAs {ncta}; per-CTA cycles (median / p10 / p90):")
for k, name in enumerate(["prefetch issue", "phase 1 (vertical sums)", "phase 2 (horizontal + solve)", "phase 3 (tail)"]):
    lines.append(f"  {name:30s} {np.median(d[:, k]):9.0f} {np.percentile(d[:, k], 10):9.0f} {np.percentile(d[:, k], 90):9.0f}")
life = (t[:, 4] - t[:, 0]).astype(np.int64)
lines.append(f"  {'CTA lifetime':30s} {np.median(life):9.0f} {np.percentile(life, 10):9.0f} {np.percentile(life, 90):9.0f}")
# overlap on one SM: for each SM, fraction of time with k CTAs in each phase
sm = t[:, 7].astype(np.int64)
gt = t[:, 6].astype(np.int64)
lines.append(f"kernel span by globaltimer: {(gt.max() - gt.min()) / 1e3:.1f} us between first and last CTA start")
sel = np.flatnonzero(sm == sm[0])
ev = []
for i in sel:
    for k in range(4):
        ev.append((int(t[i, k]), k, +1)); ev.append((int(t[i, k + 1]), k, -1))
ev.sort()
occ = [0, 0, 0, 0]; last = ev[0][0]; acc = {}
for tt, k, dlt in ev:
    key = tuple(occ); acc[key] = acc.get(key, 0) + (tt - last); last = tt
    occ[k] += dlt
tot = sum(acc.values())
lines.append(f"SM {sm[0]}: {len(sel)} CTAs; share of time by (n in prefetch, n in P1, n in P2, n in P3):")
for key, v in sorted(acc.items(), key=lambda kv: -kv[1])[:14]:
    lines.append(f"  {key}  {100 * v / tot:5.1f} %")
txt = "\n".join(lines)
print(txt)
(out / "trace_phases.txt").write_text(txt + "\n")
