#!/usr/bin/env python
"""Headline benchmark: Farneback flow -> PC1 frame-pairs/s at 1080p (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1|C2|C4|C5]
    python bench.py --scaling strong [--clips 64 --frames 300]            # config C3: a fixed batch of clips sharded by clip
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the whole hot path over one batch of synthetic input per GPU: `pairs_per_step` frame
pairs of the configured clip (default C2: 1920x1080, cv2 default Farneback parameters, full-frame ROI): pyramid ->
polynomial expansion -> [update matrices + blur + solve] x iterations x scales -> body-axis projection + ROI means,
gather of the per-frame series to rank 0, NaN-robust zero-phase band-pass, sliding-window PCA -> PC1 (all CUDA).
  value   : whole-job pairs/s with the frames already resident in HBM (device timed, max over ranks)
  e2e     : same through the host-buffer API (pinned host frames -> H2D inside the timed region -> series/PC1 D2H)
  roofline: every stage of the flow call bracketed by CUDA events on its launch stream (bf_plan_profile).  `frac` is the
            dominant kernel -- fused blur + solve + UpdateMatrices at the finest scale -- on its algorithmic 56 B per pixel
            (flow r/w 8+8, R0 20, gathered R1 20); the last iteration of a scale (no update) is reported
            separately, `stage_frac` covers the whole finest scale (first update + all iterations) on I x 56 B per pixel
            and `pipeline_frac` the whole step on BASELINE's bytes per pair.
  parity  : measured, not only asserted in tests: the benchmarked plan (and the exact plan) against cv2 on frames of this
            clip -- dense field of two pairs taken from a full batch, ROI rows over a longer run, PC1 correlation.
  cpu_baseline / --impl reference: the reference's own CPU path (cv2.calcOpticalFlowFarneback + numpy reduction,
            through oracle/cv2_ref.py) on this box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "farneback_flow_to_pc1_frame_pairs_per_s_1080p"
UNIT = "frame-pairs/s"
FALLBACK_HBM_GBS = 6650.0
CONFIG_TEXT = {
    "C1": "C1: synthetic 640x480 30 fps clip, decaying 3 Hz chirp patch, 200x200 ROI, cv2 default Farneback params",
    "C2": "C2: synthetic 1920x1080 30 fps VEEG-like clip (decaying 3 Hz chirp patch), full-frame ROI, cv2 default Farneback params",
    "C3": "C3: batch of synthetic 1080p clips (C2 parameters, one seed per clip) sharded by clip across the GPUs",
    "C4": "C4: synthetic 3840x2160 60 fps clip, OPTFLOW_FARNEBACK_GAUSSIAN, poly_n 7, sigma 1.5, winsize 21, levels 5",
    "C5": "C5: synthetic 1920x1080 clip, 2 ROIs (bilateral), short-time PCA window sweep 0.5/1/2/4 s",
}


def algorithmic_bytes_per_pair(W, H, S, iters):
    """SURVEY 8d / BASELINE.md section 3, streaming mode: W*H + S*(28 + 56*I)."""
    return W * H + S * (28 + 56 * iters)


def measured_traffic():
    """DRAM bytes per pair-iteration of the two variants of the dominant kernel, from the committed ncu --set full capture at
    the bench's own 64 pairs per launch (profiles/r2_dominant_kernel.json; command inside)."""
    try:
        return json.loads((ROOT / "profiles" / "r2_dominant_kernel.json").read_text())
    except Exception:
        return None


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int):
        self.samples, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1])); pw.append(float(s[2]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, s[3:7]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---- workload description ------------------------------------------------------------------------------------------
def pca_configs(cfg: str, fps: float):
    """(win_sec list, win_n list, step_n list): C5 sweeps the short-time PCA window, everything else uses the reference's 2 s."""
    from oracle import pc1_np
    wins = [0.5, 1.0, 2.0, 4.0] if cfg == "C5" else [2.0]
    ws = [pc1_np.window_samples(w, 0.1, fps) for w in wins]
    return wins, [w for w, _ in ws], [s for _, s in ws]


def workload_config(cfg, spec, params, pairs_per_step, batch, n_roi, wins, storage=None):
    return {"workload": CONFIG_TEXT[cfg] + ", flow->ROI series->band-pass->PC1",
            "storage": storage, "width": spec.W, "height": spec.H, "fb_params": params,
            "pairs_per_step_per_gpu": pairs_per_step, "pairs_per_launch": batch, "n_roi": n_roi,
            "pca": {"win_sec": wins, "step_sec": 0.1, "fs": spec.fps},
            "l2_policy": "inputs larger than L2 (each step streams >= 0.5 GB of distinct frames per GPU)",
            "parallelism": "temporal frame-chunk sharding, 1-frame overlap, series gathered to rank 0"}


def storage_text(bits):
    if bits == 16:
        return ("compact plan: polynomial coefficients as 16-byte pixels (b fp32, A fp16), update matrices as fp16 G + fp32 h "
                "with consistent rounding; storage only, all arithmetic fp32; parity measured in this line (`parity`)")
    return "exact plan: polynomial coefficients and update matrices as fp32 planes"


# ---- CPU reference legs (oracle/cv2_ref.py: the reference's own cv2 call + numpy reduction) ---------------------------
def cpu_reference_pass(frames, masks, params, fps, wins, threads=None):
    """One pass of the reference CPU path over `frames` [n+1, H, W]: cv2 flow + ROI means for every pair (ThreadPool x
    cv2.setNumThreads(1)), host band-pass (scipy) and the sequential sliding PCA.  Returns rows [n_roi, n+1, 3], PC1
    [n_cfg, n_roi, n+1], seconds."""
    from btcs_pnes_optical_flow_b200 import pca
    from oracle import cv2_ref, pc1_np
    threads = threads or os.cpu_count() or 1
    sos = pca.butter_bandpass_sos(pca.BPF_LOW_HZ, pca.BPF_HIGH_HZ, fps)
    t0 = time.perf_counter()
    rows = cv2_ref.roi_series(frames, [1.0, 0.0], [0.0, 1.0], masks, params, threads=threads)
    pc1 = np.full((len(wins), rows.shape[0], rows.shape[1]), np.nan)
    for r in range(rows.shape[0]):
        bx, by = pca.bandpass_nanrobust(rows[r, :, 0], sos), pca.bandpass_nanrobust(rows[r, :, 1], sos)
        for c, w in enumerate(wins):
            win_n, step_n = pc1_np.window_samples(w, 0.1, fps)
            pc1[c, r] = pc1_np.dynamic_pc1_sliding(bx, by, win_n, step_n)
    return rows, pc1, time.perf_counter() - t0, threads


def cpu_default_threading(frames, masks, params, n_pairs=3):
    """the reference path exactly as the script runs it -- one pair after the other, cv2's
    default threading (cv2.getNumThreads() = core count; cv2's Farneback itself barely uses them)."""
    import cv2
    from oracle import cv2_ref
    n = min(n_pairs, frames.shape[0] - 1)
    t0 = time.perf_counter()
    for t in range(1, n + 1):
        cv2_ref.roi_mean_body_flow(frames[t - 1], frames[t], [1.0, 0.0], [0.0, 1.0], masks[0], params)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "pairs": n, "cv2_threads": int(cv2.getNumThreads())}


def run_reference(args, cfg, spec, params):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    from oracle import cv2_ref
    cores = os.cpu_count() or 1
    per_pair_s = 0.7 * (spec.W * spec.H) / (1920 * 1080)            # one core, cv2 defaults; sized so a step stays bounded
    n_pairs = int(max(4, min(2 * cores, 64, 20.0 * cores / max(per_pair_s, 1e-3))))
    wins, _, _ = pca_configs(cfg, spec.fps)
    frames = syn.make_clip_np(spec, 0, n_pairs + 1)
    masks = spec.roi_masks()
    for _ in range(args.warmup):
        cpu_reference_pass(frames, masks, params, spec.fps, wins)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, _, threads = cpu_reference_pass(frames, masks, params, spec.fps, wins)
    dt = time.perf_counter() - t0
    value = n_pairs * args.steps / dt
    sample = (f"{n_pairs} consecutive {spec.W}x{spec.H} frame pairs per step of the same synthetic clip; cv2 {cv2_ref.cv2_version()}, "
              f"ThreadPool({threads}) x cv2.setNumThreads(1)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, spec, params, n_pairs, None, masks.shape[0], wins),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
                         "default_threading": cpu_default_threading(frames, masks, params)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- roofline / parity blocks ----------------------------------------------------------------------------------------
def roofline_block(prof, ms_total, fine, iters, peak, peak_src, value_per_gpu, bpp, kernel_name):
    st = prof["stages"]
    px = fine["w"] * fine["h"]
    gbs = lambda pairs, nbytes, ms: (pairs * nbytes * px / (ms / 1e3) / 1e9) if ms > 0 else None
    upd, last, first = st["iter_update"], st["iter_last"], st["update"]
    ach = gbs(upd["pairs"], 56, upd["ms"])
    traffic = measured_traffic()
    # the capture is of the 1080p box-window kernel at 64 pairs per launch: other frame sizes, windows or batches get null
    if traffic and not ((fine["w"], fine["h"]) == (1920, 1080) and "box<7" in kernel_name and "compact" in kernel_name
                        and upd["launches"] and upd["pairs"] == traffic.get("pairs_per_launch", 64) * upd["launches"]):
        traffic = None
    block = {
        "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
        "traffic": (traffic["iter_update_dram_bytes_per_pair"] * upd["pairs"] / max(upd["launches"], 1)) if traffic else None,
        "kernel": kernel_name + ": blur + solve + UpdateMatrices launches at the finest scale (the launches that read M, R0, "
                  "gathered R1 and write M')",
        "launches": upd["launches"], "avg_launch_ms": upd["ms"] / max(upd["launches"], 1),
        "algorithmic_bytes_per_launch": 56 * px * upd["pairs"] / max(upd["launches"], 1),
        "kernel_share_of_step": upd["ms"] / ms_total, "peak_source": peak_src,
        "last_iteration": {"launches": last["launches"], "avg_launch_ms": last["ms"] / max(last["launches"], 1),
                           "share_of_step": last["ms"] / ms_total,
                           "note": "same kernel without the update tail (reads M only; ROI sums or coarse flow out): not credited "
                                   "with 56 B/px, reported on its own",
                           "traffic": (traffic["iter_last_dram_bytes_per_pair"] * last["pairs"] / max(last["launches"], 1)) if traffic else None},
        "stage_frac": None, "pipeline_bytes_per_pair": bpp, "pipeline_achieved_GBs_per_gpu": value_per_gpu * bpp / 1e9,
        "pipeline_frac": value_per_gpu * bpp / 1e9 / peak,
        "stages_ms_timed_region": {k: v["ms"] for k, v in st.items()}, "flow_stages_share_of_step": sum(v["ms"] for v in st.values()) / ms_total,
        "traffic_source": traffic.get("source") if traffic else None,
    }
    stage_ms = first["ms"] + upd["ms"] + last["ms"]
    if stage_ms > 0 and first["pairs"]:
        stage_gbs = first["pairs"] * iters * 56 * px / (stage_ms / 1e3) / 1e9
        block["stage_frac"] = stage_gbs / peak
        block["stage"] = {"what": "finest scale: first UpdateMatrices (+ flow upsample) + all iterations, on iterations x 56 B/px",
                          "achieved": stage_gbs, "ms_timed_region": stage_ms, "share_of_step": stage_ms / ms_total}
    return block


def flow_metrics(got, ref, band):
    d = np.sqrt(((np.asarray(got, np.float64) - np.asarray(ref, np.float64)) ** 2).sum(-1))
    inner = d[band:-band, band:-band]
    edge = d.copy()
    edge[band:-band, band:-band] = 0.0
    return {"mean_epe": float(d.mean()), "interior_max": float(inner.max()), "band_max": float(edge.max()),
            "frac_gt_0.05": float((d > 0.05).mean()), "frac_gt_0.01": float((d > 0.01).mean())}


def parity_block(B, spec, params, frames_dev, masks, wins, max_pairs, n_par, local_rank, cfg):
    """Both storage modes against cv2 on frames of the benchmarked clip, outside the timed region (SURVEY 8d gates: mean EPE
    <= 0.01 px, max <= 0.05 px on the flow field, PC1 r >= 0.9999).  Also yields the CPU baseline (same pass)."""
    import torch
    from btcs_pnes_optical_flow_b200 import pca
    from oracle import cv2_ref
    dev = frames_dev.device
    n_par = min(n_par, frames_dev.shape[0] - 1)
    n_dense = min(max_pairs, n_par)
    frames_h = frames_dev[:n_par + 1].cpu().numpy()
    cv_rows, cv_pc1, cpu_s, threads = cpu_reference_pass(frames_h, masks, params, spec.fps, wins)
    dense_idx = sorted({min(5, n_dense - 1), n_dense - 2 if n_dense >= 2 else 0})
    cv_dense = {k: cv2_ref.farneback(frames_h[k], frames_h[k + 1], **params) for k in dense_idx}
    band = 2 * (params["winsize"] // 2) + 2
    masks_dev = torch.from_numpy(masks.astype(np.uint8)).to(dev)
    sos = pca.butter_bandpass_sos(pca.BPF_LOW_HZ, pca.BPF_HIGH_HZ, spec.fps)
    _, win_n, step_n = pca_configs(cfg, spec.fps)
    out = {"cv2": cv2_ref.cv2_version(), "pairs_rows": n_par, "pairs_per_launch": n_dense, "dense_pairs": dense_idx, "band_px": band,
           "gates": {"mean_epe": 0.01, "max_epe": 0.05, "pc1_r": 0.9999},
           "note": "band = outer band_px pixels: on a static border the reference's `inside` branch flips on the sign of a "
                   "numerically-zero flow, which no non-bit-identical build reproduces (tests/test_oracle_farneback.py::"
                   "test_static_border_branch_flip_is_inherent: the NumPy restatement vs cv2 shows the same band)"}
    for kind in ("compact", "exact"):
        with B.FlowPlan(spec.W, spec.H, params, max_pairs=n_dense if kind == "compact" else min(n_dense, 16), max_rois=masks.shape[0],
                        device=local_rank, exact=(kind == "exact")) as plan:
            _, flow = plan.flow_series(frames_dev[:n_dense + 1], None, None, masks_dev, return_flow=True)
            dense = [flow_metrics(flow[k].cpu().numpy(), cv_dense[k], band) for k in dense_idx]
            del flow
            rows = plan.flow_series(frames_dev[:n_par + 1], None, None, masks_dev)           # ring wrap when n_par > max_pairs
            s = rows.double()
            both = pca.bandpass_nanrobust_device(torch.cat([s[:, :, 0], s[:, :, 1]]), sos)
            nr = masks.shape[0]
            pc1 = pca.pc1_sliding_batched(both[:nr].contiguous(), both[nr:].contiguous(), win_n, step_n).cpu().numpy()
            rows = rows.cpu().numpy()
        ok = np.isfinite(cv_rows) & np.isfinite(rows)
        assert ok.any() and np.array_equal(np.isfinite(cv_rows), np.isfinite(rows)), "NaN rows differ from the reference's"
        rs = []
        for c in range(len(wins)):
            for r in range(nr):
                m = np.isfinite(pc1[c, r]) & np.isfinite(cv_pc1[c, r])
                if m.sum() >= 8 and np.std(cv_pc1[c, r][m]) > 0:
                    rs.append(float(np.corrcoef(pc1[c, r][m], cv_pc1[c, r][m])[0, 1]))
        res = {k: max(d[k] for d in dense) for k in dense[0]}
        res.update(roi_mean_abs_err=float(np.abs(rows[ok] - cv_rows[ok]).max()), pc1_r=min(rs) if rs else None)
        out[kind] = res
    cpu = {"value": n_par / cpu_s, "unit": UNIT, "cores": threads, "kind": "reference",
           "sample": f"{n_par} consecutive {spec.W}x{spec.H} pairs of the same clip, 1 pass ({cpu_s:.1f} s); cv2 {cv2_ref.cv2_version()} "
                     f"calcOpticalFlowFarneback + numpy ROI reduction + scipy band-pass + sliding PCA, ThreadPool({threads}) x "
                     f"cv2.setNumThreads(1)",
           "default_threading": cpu_default_threading(frames_h, masks, params)}
    return out, cpu


# ---- our arm -----------------------------------------------------------------------------------------------------------
def setup_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def run_ours(args, cfg, spec, params):
    import torch
    import torch.distributed as dist
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import _lib, distributed as D, pca, synthetic as syn

    world, rank, local_rank, dev = setup_dist()
    lib = _lib.load()
    P = args.pairs_per_step
    T_total = world * P + 1
    rows = D.shard_rows(T_total, world)
    lo, hi = rows[rank]
    f0, f1 = D.frames_for_rows(lo, hi)
    spec.T = T_total
    frames_dev = syn.make_clip(spec, dev, f0, f1 - f0)                       # uint8 [P+1, H, W] resident in HBM
    masks_host = spec.roi_masks()
    n_roi = masks_host.shape[0]
    masks_dev = torch.from_numpy(masks_host).to(dev)
    frames_host = frames_dev.cpu().pin_memory().numpy()
    plan = B.FlowPlan(spec.W, spec.H, params, max_pairs=args.max_pairs, max_rois=n_roi, device=local_rank)
    sos = pca.butter_bandpass_sos(pca.BPF_LOW_HZ, pca.BPF_HIGH_HZ, spec.fps)
    wins, win_n, step_n = pca_configs(cfg, spec.fps)

    side = torch.cuda.Stream(device=dev)        # the PC1 tail runs beside the next step's flow kernels, not behind them

    def finish(full, ready=None):
        """rank 0: series -> band-pass -> sliding PCA -> PC1 [n_cfg, n_roi, T], all on the GPU; only PC1 comes back."""
        if ready is not None:
            side.wait_event(ready)
        with torch.cuda.stream(side):
            if not full.is_cuda:
                full = full.to(dev)
            else:
                full.record_stream(side)
            s = full.double()
            both = pca.bandpass_nanrobust_device(torch.cat([s[:, :, 0], s[:, :, 1]]), sos)
            return pca.pc1_sliding_batched(both[:n_roi].contiguous(), both[n_roi:].contiguous(), win_n, step_n).cpu().numpy()

    def make_steps(the_plan):
        def launch_device():
            """Asynchronous part of a step: flow -> ROI series on the device, gather to rank 0 (NCCL)."""
            series = the_plan.flow_series(frames_dev, None, None, masks_dev)        # [n_roi, P+1, 3] on the device
            full = D.gather_series(series[:, 1:], rows, T_total)
            ready = torch.cuda.Event()
            ready.record()
            return full, ready

        def tail_device(launched):
            full, ready = launched
            return finish(full, ready) if rank == 0 else None
        return launch_device, tail_device

    launch_device, tail_device = make_steps(plan)

    def launch_host():
        """Streaming host-buffer call: pinned frames in, queued H2D chunks + compute, series D2H; returns a handle."""
        return plan.flow_series_async(frames_host, None, None, masks_host)

    def tail_host(handle):
        series = handle.result()                                               # this step's series has reached the host
        if world > 1:
            with torch.cuda.stream(side):                                      # the gather must not queue behind the next step
                full = D.gather_series(torch.from_numpy(series[:, 1:]).to(dev), rows, T_total)
        else:
            full = torch.from_numpy(series)
        return finish(full) if rank == 0 else None

    def timed(steps, warmup, launch, tail, prof_plan=None):
        """The tail of step i (series gather, band-pass, PC1, D2H) runs after step i+1 has been queued, as a streaming caller
        would do it; the work per step is the same."""
        for _ in range(warmup):
            tail(launch())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.bf_launch_count_reset()
        if prof_plan is not None:
            prof_plan.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, pending = None, None
        for _ in range(steps):
            nxt = launch()
            if pending is not None:
                out = tail(pending)
            pending = nxt
        out = tail(pending)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launches = int(lib.bf_launch_count())
        prof = None
        if prof_plan is not None:
            prof_plan.profile(False)
            prof = prof_plan.profile_read()
        return float(ms.item()), launches, prof, out

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, launches, prof, pc1 = timed(args.steps, args.warmup, launch_device, tail_device, prof_plan=plan)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _, _, pc1_h = timed(args.steps, args.warmup, launch_host, tail_host)

    if rank == 0:
        pairs = world * P * args.steps
        value = pairs / (ms_total / 1e3)
        e2e_value = pairs / (ms_e2e / 1e3)
        sc = plan.scales()
        S = sum(s["w"] * s["h"] for s in sc)
        iters = params["iterations"]
        bpp = algorithmic_bytes_per_pair(spec.W, spec.H, S, iters)
        peak, peak_src = hbm_peak()
        kname = ("k_blur_solve_gauss" if params["flags"] & 256 else "k_blur_solve_box") + f"<{params['winsize'] // 2}, compact>"
        roof = roofline_block(prof, ms_total, sc[-1], iters, peak, peak_src, value / world, bpp, kname) if prof and prof["launches"] else None
        assert pc1 is not None and pc1.shape == (len(wins), n_roi, T_total)
        both = np.isfinite(pc1) & np.isfinite(pc1_h)
        if T_total >= 2 * max(win_n):           # long enough for the PCA windows: the series must be usable
            assert both[wins.index(2.0) if 2.0 in wins else 0, 0].sum() > 0.9 * (T_total - 1), "PC1 mostly NaN"
            assert np.corrcoef(pc1[both], pc1_h[both])[0, 1] > 0.999999, "host-buffer and device-buffer paths disagree"
        parity, cpu = None, None
        if world == 1 and not args.no_parity:
            parity, cpu = parity_block(B, spec, params, frames_dev, masks_host, wins, args.max_pairs, args.parity_pairs, local_rank, cfg)
        exact = None
        if world == 1 and not args.no_exact:
            # the same device-resident step on an exact (all-fp32 storage) plan, with its own stage profile
            plan_x = B.FlowPlan(spec.W, spec.H, params, max_pairs=args.max_pairs, max_rois=n_roi, device=local_rank, exact=True)
            lx, tx = make_steps(plan_x)
            ms_x, _, prof_x, _ = timed(2, 2, lx, tx, prof_plan=plan_x)
            vx = P * 2 / (ms_x / 1e3)
            exact = {"value": vx, "unit": UNIT, "storage": storage_text(32), "steps": 2, "pairs_per_launch": args.max_pairs,
                     "roofline": roofline_block(prof_x, ms_x, sc[-1], iters, peak, peak_src, vx, bpp, kname.replace("compact", "exact"))}
            if exact["roofline"]:
                exact["roofline"]["traffic"] = None
                exact["roofline"]["last_iteration"]["traffic"] = None
            plan_x.close()
        h2d = frames_host.nbytes + masks_host.size + 2 * (P + 1) * 2 * 8
        d2h = n_roi * (P + 1) * 3 * 4 + len(wins) * n_roi * T_total * 8
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, spec, params, P, args.max_pairs, n_roi, wins, storage_text(plan.coeff_storage_bits)),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "workspace_bytes": plan.workspace_bytes, "exact_f32_storage": exact,
        }
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def run_strong(args, spec, params):
    """Config C3: a FIXED batch of clips (--clips x --frames) sharded by clip over the ranks: each rank runs the
    whole path for its clips -- flow -> series -> band-pass -> PC1 on the owning GPU -- and only PC1 [clips, frames] is
    gathered to rank 0.  Total work does not grow with N: "scaling": "strong"."""
    import torch
    import torch.distributed as dist
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import _lib, distributed as D, pca, synthetic as syn

    world, rank, local_rank, dev = setup_dist()
    lib = _lib.load()
    n_clips, F = args.clips, args.frames
    mine = D.shard_clips(n_clips, world)[rank]
    most = max(len(c) for c in D.shard_clips(n_clips, world))
    spec.T = F
    masks_host = spec.roi_masks()
    n_roi = masks_host.shape[0]
    masks_dev = torch.from_numpy(masks_host).to(dev)
    clips_dev = []
    for c in mine:
        spec.seed = 1000 + c                                                   # every clip its own texture and phase
        clips_dev.append(syn.make_clip(spec, dev, 0, F))
    distinct = min(len(mine), 4)                                               # host copies: a few distinct clips, cycled (same bytes cross PCIe)
    clips_host = [clips_dev[i].cpu().pin_memory().numpy() for i in range(distinct)]
    plan = B.FlowPlan(spec.W, spec.H, params, max_pairs=args.max_pairs, max_rois=n_roi, device=local_rank)
    sos = pca.butter_bandpass_sos(pca.BPF_LOW_HZ, pca.BPF_HIGH_HZ, spec.fps)
    wins, win_n, step_n = pca_configs("C3", spec.fps)

    def pc1_of(series):                                                        # [n_local, n_roi, F, 3] on the device -> PC1 [n_local, F]
        s = series[:, 0].double()
        both = pca.bandpass_nanrobust_device(torch.cat([s[:, :, 0], s[:, :, 1]]), sos)
        n = s.shape[0]
        return pca.pc1_sliding_batched(both[:n].contiguous(), both[n:].contiguous(), win_n, step_n)[0]

    def gather_pc1(local):
        pad = torch.full((most, F), float("nan"), dtype=torch.float64, device=dev)
        pad[:local.shape[0]] = local
        if world == 1:
            return pad.cpu().numpy()
        parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, parts, dst=0)
        return torch.stack(parts).cpu().numpy() if rank == 0 else None

    def step_device():
        series = torch.stack([plan.flow_series(cl, None, None, masks_dev) for cl in clips_dev]) if clips_dev else \
            torch.empty((0, n_roi, F, 3), device=dev)
        return gather_pc1(pc1_of(series) if clips_dev else torch.empty((0, F), dtype=torch.float64, device=dev))

    def step_host():
        pend = [plan.flow_series_async(clips_host[i % distinct], None, None, masks_host) for i in range(len(mine))]
        series = torch.stack([torch.from_numpy(p.result()).to(dev) for p in pend]) if pend else torch.empty((0, n_roi, F, 3), device=dev)
        return gather_pc1(pc1_of(series) if pend else torch.empty((0, F), dtype=torch.float64, device=dev))

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.bf_launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), int(lib.bf_launch_count()), out

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, launches, pc1 = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _, _ = timed(step_host, args.steps, 1)
    if rank == 0:
        total_pairs = n_clips * (F - 1)
        value = total_pairs * args.steps / (ms_total / 1e3)
        sc = plan.scales()
        bpp = algorithmic_bytes_per_pair(spec.W, spec.H, sum(s["w"] * s["h"] for s in sc), params["iterations"])
        peak, peak_src = hbm_peak()
        assert np.isfinite(pc1).mean() > 0.9
        cfgd = workload_config("C3", spec, params, len(mine) * (F - 1), args.max_pairs, n_roi, wins, storage_text(plan.coeff_storage_bits))
        cfgd.update(clips=n_clips, frames_per_clip=F, clips_per_gpu=[len(c) for c in D.shard_clips(n_clips, world)],
                    parallelism="clips dealt round-robin to ranks; per-clip PC1 on the owning rank; one gather of PC1 [clips, frames] to rank 0",
                    e2e_host_copies=f"{distinct} distinct pinned clips per rank, cycled over its {len(mine)} clips (same H2D volume)")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfgd, "clocks": clocks,
            "e2e": {"value": total_pairs * args.steps / (ms_e2e / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": int(len(mine) * (F * spec.W * spec.H + masks_host.size + 2 * F * 16)),
                    "d2h_bytes_per_step": int(len(mine) * n_roi * F * 12 + n_clips * F * 8), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "pipeline_bytes_per_pair": bpp,
                         "achieved": value / world * bpp / 1e9, "frac": value / world * bpp / 1e9 / peak,
                         "pipeline_frac": value / world * bpp / 1e9 / peak, "traffic": None,
                         "note": "whole-pipeline figure per GPU (per-kernel numbers: the default weak-scaling line)"},
            "cpu_baseline": None, "workspace_bytes": plan.workspace_bytes,
        }
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C4", "C5"], help="BASELINE.json configuration (default: the headline C2)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: pairs_per_step pairs per GPU; strong: config C3, a fixed batch of --clips clips sharded by clip")
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--pairs-per-step", type=int, default=None, help="frame pairs per GPU per step (default 256; C1 1024; C4 32)")
    ap.add_argument("--max-pairs", type=int, default=None, help="frame pairs per batched kernel launch (default 64; C4 16)")
    ap.add_argument("--parity-pairs", type=int, default=None, help="pairs of the clip compared with cv2 outside the timed region")
    ap.add_argument("--no-exact", action="store_true", help="skip the secondary measurement on an exact (fp32 storage) plan")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity / CPU-baseline block")
    args = ap.parse_args()
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    cfg = "C3" if args.scaling == "strong" else args.config
    spec, params = syn.config_spec(cfg)
    defaults = {"C1": (1024, 64, 128), "C2": (256, 64, 96), "C3": (0, 64, 0), "C4": (32, 16, 6), "C5": (256, 64, 96)}[cfg]
    args.pairs_per_step = args.pairs_per_step or defaults[0]
    args.max_pairs = args.max_pairs or defaults[1]
    args.parity_pairs = args.parity_pairs or defaults[2]
    if args.impl == "reference":
        run_reference(args, cfg, spec, params)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), __file__,
                   *sys.argv[1:]]
            raise SystemExit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.scaling == "strong":
        run_strong(args, spec, params)
    else:
        run_ours(args, cfg, spec, params)


if __name__ == "__main__":
    main()
