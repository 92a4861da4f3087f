"""Temporal data-parallel sharding of the flow stage.

Pair t needs frames t-1 and t only (/root/reference/optical_flow.py:242-249) and clips are independent, so a
clip is cut into contiguous frame chunks that overlap by ONE frame, one chunk per rank (one process per GPU),
with no data-path collective.  The only exchange is a gather of the per-frame ROI series (KBs) to rank 0, where
the host band-pass and the PC1 stage run.  Works with any torch.distributed backend: NCCL over NVLink on the
B200 box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_rows(T: int, world: int) -> list[tuple[int, int]]:
    """Split output rows 1..T-1 (row t = pair (t-1, t)) into `world` contiguous, balanced [lo, hi) ranges."""
    n = max(T - 1, 0)
    base, rem = divmod(n, world)
    out, lo = [], 1
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        out.append((lo, lo + cnt))
        lo += cnt
    return out


def frames_for_rows(lo: int, hi: int) -> tuple[int, int]:
    """Frame range [f0, f1) a rank must hold to produce rows [lo, hi): one extra frame in front."""
    return (lo - 1, hi) if hi > lo else (lo, lo)


def shard_clips(n_clips: int, world: int) -> list[list[int]]:
    """Config C3: whole clips dealt round-robin to ranks."""
    return [list(range(r, n_clips, world)) for r in range(world)]


def gather_series(local: torch.Tensor, rows: Sequence[tuple[int, int]], T: int, dst: int = 0,
                  group=None) -> torch.Tensor | None:
    """Gather per-rank series chunks [n_roi, hi-lo, 3] into the full [n_roi, T, 3] on rank `dst` (row 0 = NaN).

    Chunks can differ by one row, so they are padded to the longest; one collective of a few KB."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_roi = local.shape[0]
    longest = max(hi - lo for lo, hi in rows)
    pad = torch.full((n_roi, longest, 3), float("nan"), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    if world == 1:
        parts = [pad]
    else:
        parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, parts, dst=dst, group=group)
        if rank != dst:
            return None
    full = torch.full((n_roi, T, 3), float("nan"), dtype=local.dtype, device=local.device)
    for (lo, hi), part in zip(rows, parts):
        full[:, lo:hi] = part[:, :hi - lo]
    return full


def sharded_flow_series(compute_chunk: Callable[[int, int], torch.Tensor], T: int, dst: int = 0,
                        group=None) -> torch.Tensor | None:
    """Run `compute_chunk(f0, f1)` -> [n_roi, f1-f0, 3] (row 0 of the chunk is the NaN row of its first frame)
    on this rank's frame range and gather the rows to `dst`.  `compute_chunk` is the GPU flow series in
    production (FlowPlan.flow_series on frames[f0:f1]) and a CPU stand-in in the gloo tests."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rows = shard_rows(T, world)
    lo, hi = rows[rank]
    f0, f1 = frames_for_rows(lo, hi)
    local = None
    if hi > lo:
        chunk = compute_chunk(f0, f1)
        local = chunk[:, 1:]                      # drop the chunk's own NaN row (its first frame has no prev here)
    if world > 1:
        # A rank without rows (more ranks than pairs) never calls compute_chunk -- in production that would be a flow
        # series over zero frames -- and takes the shape of its empty chunk from rank 0, which owns rows whenever T > 1.
        meta = [None]
        if rank == 0 and local is not None:
            meta = [(int(local.shape[0]), str(local.dtype).split(".")[-1])]
        dist.broadcast_object_list(meta, src=0, group=group)
        if local is None:
            n_roi, dtype = meta[0] if meta[0] is not None else (1, "float32")
            backend = dist.get_backend(group)
            dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
            local = torch.empty((n_roi, 0, 3), dtype=getattr(torch, dtype), device=dev)
    elif local is None:
        local = torch.empty((1, 0, 3), dtype=torch.float32)
    return gather_series(local, rows, T, dst, group)
