#!/usr/bin/env python
"""Headline benchmark: Farneback flow -> PC1 frame-pairs/s at 1080p (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the whole hot path over one batch of synthetic input per GPU: `pairs_per_step` frame
pairs of the config-C2 clip (1920x1080, cv2 default Farneback parameters, full-frame ROI): pyramid -> polynomial
expansion -> [update matrices + blur + solve] x iterations x scales -> body-axis projection + ROI means, gather of
the per-frame series to rank 0, NaN-robust zero-phase band-pass, sliding-window PCA -> PC1 (all CUDA).
  value  : whole-job pairs/s with the frames already resident in HBM (device timed, max over ranks)
  e2e    : same through the host-buffer API (pinned host frames -> H2D inside the timed region -> series/PC1 D2H)
  roofline: the dominant kernel (fused blur+solve[+update] at the finest scale) timed with CUDA events on its
           launch stream; algorithmic bytes = 56 B per pixel per pair-iteration
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


def algorithmic_bytes_per_pair(W, H, S, iters):
    """SURVEY 8d / BASELINE.md section 3, streaming mode: W*H + S*(28 + 56*I)."""
    return W * H + S * (28 + 56 * iters)


def measured_traffic_per_pair_iteration():
    """DRAM bytes per pair-iteration of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    p = ROOT / "profiles" / "r1_dominant_kernel.json"
    try:
        return float(json.loads(p.read_text())["dram_bytes_per_pair_iteration"])
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


def cpu_reference_run(spec, params, n_pairs, steps, warmup, threads=None):
    """The reference's CPU path (flow -> ROI series -> band-pass -> PC1) on `n_pairs` pairs per step."""
    from btcs_pnes_optical_flow_b200 import pca, synthetic as syn
    from oracle import cv2_ref, pc1_np
    threads = threads or os.cpu_count() or 1
    frames = syn.make_clip_np(spec, 0, n_pairs + 1)
    mask = spec.roi_mask()
    t = np.arange(n_pairs + 1) / spec.fps
    sos = pca.butter_bandpass_sos(pca.BPF_LOW_HZ, pca.BPF_HIGH_HZ, spec.fps)
    win_n, step_n = pc1_np.window_samples(pca.WIN_SEC, pca.STEP_SEC, spec.fps)

    def step():
        rows = cv2_ref.roi_series(frames, [1.0, 0.0], [0.0, 1.0], mask, params, threads=threads)[0]
        return pc1_np.dynamic_pc1_sliding(pca.bandpass_nanrobust(rows[:, 0], sos), pca.bandpass_nanrobust(rows[:, 1], sos),
                                          win_n, step_n)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_pairs * steps / dt, dt / steps * 1e3, threads, cv2_ref.cv2_version()


def run_reference(args, spec, params):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_pairs = max(8, min(2 * cores, 64))
    value, ms, threads, ver = cpu_reference_run(spec, params, n_pairs, args.steps, args.warmup)
    sample = f"{n_pairs} consecutive 1080p frame pairs per step of the same synthetic clip; cv2 {ver}, " \
             f"ThreadPool({threads}) x cv2.setNumThreads(1)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(spec, params, n_pairs, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(spec, params, pairs_per_step, batch, storage_bits=None):
    return {"storage": None if storage_bits is None else
            (f"polynomial coefficients and update matrices stored as fp{storage_bits} "
             f"({'compact plan, storage only: all arithmetic fp32' if storage_bits == 16 else 'exact plan'}); "
             "parity gates asserted in tests/test_gpu_flow.py"),
            "workload": f"C2: synthetic {spec.W}x{spec.H} {spec.fps:g} fps VEEG-like clip (decaying 3 Hz chirp patch), "
                        f"full-frame ROI, cv2 default Farneback params, flow->ROI series->band-pass->PC1",
            "width": spec.W, "height": spec.H, "fb_params": params, "pairs_per_step_per_gpu": pairs_per_step,
            "pairs_per_launch": batch, "pca": {"win_sec": 2.0, "step_sec": 0.1, "fs": spec.fps},
            "l2_policy": "inputs larger than L2 (each step streams >= 0.5 GB of distinct frames per GPU)",
            "parallelism": "temporal frame-chunk sharding, 1-frame overlap, series gathered to rank 0"}


def run_ours(args, spec, params):
    import torch
    import torch.distributed as dist
    import btcs_pnes_optical_flow_b200 as B
    from btcs_pnes_optical_flow_b200 import _lib, distributed as D, pca, synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    P = args.pairs_per_step
    T_total = world * P + 1
    rows = D.shard_rows(T_total, world)
    lo, hi = rows[rank]
    f0, f1 = D.frames_for_rows(lo, hi)
    spec.T = T_total
    frames_dev = syn.make_clip(spec, dev, f0, f1 - f0)                       # uint8 [P+1, H, W] resident in HBM
    mask_dev = torch.from_numpy(spec.roi_mask()).to(dev)
    frames_host = frames_dev.cpu().pin_memory().numpy()
    mask_host = spec.roi_mask()
    t_all = np.arange(T_total) / spec.fps
    plan = B.FlowPlan(spec.W, spec.H, params, max_pairs=args.max_pairs, max_rois=1, device=local_rank)
    sos = pca.butter_bandpass_sos(pca.BPF_LOW_HZ, pca.BPF_HIGH_HZ, spec.fps)

    side = torch.cuda.Stream(device=dev)        # the PC1 tail runs beside the next step's flow kernels, not behind them

    def finish(full, ready=None):
        """rank 0: series -> band-pass -> sliding PCA -> PC1, all on the GPU; only the PC1 waveform comes back."""
        if ready is not None:
            side.wait_event(ready)
        with torch.cuda.stream(side):
            if not full.is_cuda:
                full = full.to(dev)
            else:
                full.record_stream(side)
            s = full[0].double()
            return pca.flow_to_pc1(None, s[:, 0].contiguous(), s[:, 1].contiguous(), fs_hz=spec.fps, sos=sos).cpu().numpy()

    def launch_device():
        """Asynchronous part of a step: flow -> ROI series on the device, gather to rank 0 (NCCL)."""
        series = plan.flow_series(frames_dev, None, None, mask_dev)            # [1, P+1, 3] on the device
        full = D.gather_series(series[:, 1:], rows, T_total)
        ready = torch.cuda.Event()
        ready.record()
        return full, ready

    def step_device():
        return tail_device(launch_device())

    def launch_host():
        """Streaming host-buffer call: pinned frames in, queued H2D chunks + compute, series D2H; returns a handle."""
        return plan.flow_series_async(frames_host, None, None, mask_host)

    def tail_host(handle):
        series = handle.result()                                               # this step's series has reached the host
        if world > 1:
            with torch.cuda.stream(side):                                      # the gather must not queue behind the next step
                full = D.gather_series(torch.from_numpy(series[:, 1:]).to(dev), rows, T_total)
        else:
            full = torch.from_numpy(series)
        return finish(full) if rank == 0 else None

    def step_host():
        return tail_host(launch_host())

    def tail_device(launched):
        full, ready = launched
        return finish(full, ready) if rank == 0 else None

    def timed(fn, steps, warmup, profile=False, launch=None, tail=None):
        """launch=None: fn() per step.  With `launch` + `tail`, the tail of step i (series gather, band-pass, PC1, D2H)
        runs after step i+1 has been queued, as a streaming caller would do it; the work per step is the same."""
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.bf_launch_count_reset()
        if profile:
            plan.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        if launch is None:
            for _ in range(steps):
                out = fn()
        else:
            pending = None
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
        if profile:
            plan.profile(False)
            prof = plan.profile_read()
        return float(ms.item()), launches, prof, out

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, launches, prof, pc1 = timed(step_device, args.steps, args.warmup, profile=True, launch=launch_device,
                                          tail=tail_device)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _, _, pc1_h = timed(step_host, args.steps, max(1, args.warmup // 2) if args.warmup else 0, launch=launch_host,
                                tail=tail_host)

    if rank == 0:
        pairs = world * P * args.steps
        value = pairs / (ms_total / 1e3)
        e2e_value = pairs / (ms_e2e / 1e3)
        sc = plan.scales()
        S = sum(s["w"] * s["h"] for s in sc)
        iters = params["iterations"]
        bpp = algorithmic_bytes_per_pair(spec.W, spec.H, S, iters)
        peak, peak_src = hbm_peak()
        fine = sc[-1]
        roof = None
        if prof and prof["launches"]:
            bytes_per_pair_iter = 56 * fine["w"] * fine["h"]                   # flow r/w 8+8, R0 20, gathered R1 20
            achieved = prof["pair_iterations"] * bytes_per_pair_iter / (prof["total_ms"] / 1e3) / 1e9
            tpi = measured_traffic_per_pair_iteration()
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (tpi * prof["pair_iterations"] / prof["launches"]) if tpi else None,
                    "kernel": "k_blur_solve_box<7,RH,%s>: blur+solve(+update) at the finest scale" % (os.environ.get("BTCSFLOW_TILE_TH") or "16"),
                    "launches": prof["launches"], "avg_launch_ms": prof["total_ms"] / prof["launches"],
                    "algorithmic_bytes_per_launch": prof["pair_iterations"] * bytes_per_pair_iter / prof["launches"],
                    "kernel_share_of_step": prof["total_ms"] / ms_total, "peak_source": peak_src,
                    "pipeline_bytes_per_pair": bpp, "pipeline_achieved_GBs_per_gpu": value / world * bpp / 1e9,
                    "pipeline_frac": value / world * bpp / 1e9 / peak}
        assert pc1 is not None and pc1.shape == (T_total,)
        both = np.isfinite(pc1) & np.isfinite(pc1_h)
        if T_total >= 120:                      # long enough for the 2 s PCA window: the series must be usable
            assert both.sum() > 0.9 * (T_total - 1), "PC1 mostly NaN"
            assert np.corrcoef(pc1[both], pc1_h[both])[0, 1] > 0.999999, "host-buffer and device-buffer paths disagree"
        cpu = None
        if world == 1:
            cores = os.cpu_count() or 1
            n_cpu = max(16, min(2 * cores, 48))
            v, ms_cpu, threads, ver = cpu_reference_run(spec, params, n_cpu, 1, 0)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"{n_cpu} consecutive 1080p pairs of the same clip, 1 pass ({ms_cpu / 1e3:.1f} s); cv2 {ver} "
                             f"calcOpticalFlowFarneback + numpy ROI reduction, ThreadPool({threads}) x cv2.setNumThreads(1)"}
        exact = None
        if world == 1 and not args.no_exact:
            # transparency: the same device-resident step on an exact (all-fp32 storage) plan
            plan_x = B.FlowPlan(spec.W, spec.H, params, max_pairs=min(args.max_pairs, 16), max_rois=1, device=local_rank,
                                exact=True)
            def step_exact():
                series = plan_x.flow_series(frames_dev, None, None, mask_dev)
                return finish(D.gather_series(series[:, 1:], rows, T_total))
            for _ in range(2):
                step_exact()
            torch.cuda.synchronize()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            for _ in range(2):
                step_exact()
            x1.record()
            torch.cuda.synchronize()
            exact = {"value": P * 2 / (x0.elapsed_time(x1) / 1e3), "unit": UNIT, "storage": "fp32 planes (exact plan)",
                     "steps": 2}
            plan_x.close()
        n_roi = 1
        h2d = frames_host.nbytes + mask_host.size + 2 * (P + 1) * 2 * 8 + (2 * T_total * 8 if rank == 0 else 0)
        d2h = n_roi * (P + 1) * 3 * 4 + T_total * 8
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(spec, params, P, args.max_pairs, plan.coeff_storage_bits),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "workspace_bytes": plan.workspace_bytes, "exact_f32_storage": exact,
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
    ap.add_argument("--pairs-per-step", type=int, default=256, help="frame pairs per GPU per step")
    ap.add_argument("--max-pairs", type=int, default=64, help="frame pairs per batched kernel launch")
    ap.add_argument("--no-exact", action="store_true", help="skip the secondary measurement on an exact (fp32 storage) plan")
    args = ap.parse_args()
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    spec, params = syn.config_spec("C2")
    if args.impl == "reference":
        run_reference(args, spec, params)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world != args.gpus:
            if args.gpus > 1 and world == 1:
                # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
                cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                       "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), __file__,
                       *sys.argv[1:]]
                raise SystemExit(subprocess.call(cmd))
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
        run_ours(args, spec, params)


if __name__ == "__main__":
    main()
