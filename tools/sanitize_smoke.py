"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import btcs_pnes_optical_flow_b200 as B
from btcs_pnes_optical_flow_b200 import pca, synthetic as syn

def run(tag, env, params, shape=(203, 316), exact=False):
    for k, v in env.items():
        os.environ[k] = v
    spec = syn.ClipSpec(T=5, H=shape[0], W=shape[1], seed=2, patch=60, roi=80, amp=3.0)
    fr = syn.make_clip_np(spec)
    masks = np.stack([spec.roi_mask(), np.ones(shape, bool)])
    with B.FlowPlan(shape[1], shape[0], params, max_pairs=3, max_rois=2, exact=exact) as plan:
        out, flow = plan.flow_series(fr, None, None, masks, return_flow=True)
        dev = plan.flow_series(torch.from_numpy(fr).cuda(), None, None, torch.from_numpy(masks).cuda())
        torch.cuda.synchronize()
    for k in env:
        del os.environ[k]
    print(tag, "ok", float(np.nanmean(out)), flow.shape)

P = dict(B.FB_PARAMS)
G = dict(B.FB_PARAMS, levels=5, winsize=21, poly_n=7, poly_sigma=1.5, flags=256)
run("tile", {}, P)
run("tile_w4", {}, P, shape=(135, 240))
run("tile_f32", {"BTCSFLOW_R_STORAGE": "f32"}, P)
run("exact", {}, P, exact=True)
run("generic", {"BTCSFLOW_NO_FAST": "1"}, P)
run("gauss", {}, G, shape=(272, 480))
run("odd", {}, dict(B.FB_PARAMS, pyr_scale=0.7, levels=4, winsize=16, poly_n=3), shape=(131, 203))
n = 700
t = np.arange(n) / 30.0
vx = np.sin(2 * np.pi * 3 * t) * 0.6 + 0.01 * np.cos(t); vy = np.sin(2 * np.pi * 3 * t) * 0.8
vx[0] = vy[0] = np.nan; vx[100:130] = np.nan; vy[400:404] = np.nan
pc1 = pca.flow_to_pc1(t, vx, vy)
sw = pca.pc1_sliding_batched(np.stack([vx, vy]), np.stack([vy, vx]), [15, 60, 120], [3, 3, 3])
print("pc1 ok", np.isfinite(pc1).sum(), sw.shape)
