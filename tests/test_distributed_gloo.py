"""N > 1 host logic on CPU: world_size-2 gloo run of the temporal sharding + series gather.  The per-chunk compute
is the cv2 oracle standing in for the CUDA series call (same contract: [n_roi, frames, 3], row 0 NaN)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from btcs_pnes_optical_flow_b200 import distributed as D


def test_shard_rows_cover_everything_once():
    for T in (1, 2, 3, 10, 301, 9000):
        for world in (1, 2, 4, 8):
            rows = D.shard_rows(T, world)
            assert len(rows) == world
            flat = [t for lo, hi in rows for t in range(lo, hi)]
            assert flat == list(range(1, T))
            sizes = [hi - lo for lo, hi in rows]
            assert max(sizes) - min(sizes) <= 1
            for lo, hi in rows:
                f0, f1 = D.frames_for_rows(lo, hi)
                if hi > lo:
                    assert (f0, f1) == (lo - 1, hi)            # one-frame overlap with the previous chunk
    assert D.shard_clips(64, 8)[3] == list(range(3, 64, 8))
    assert sorted(sum(D.shard_clips(10, 4), [])) == list(range(10))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames, masks, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cv2_ref
        T = frames.shape[0]

        def compute_chunk(f0, f1):
            out = cv2_ref.roi_series(frames[f0:f1], [1.0, 0.0], [0.0, 1.0], masks, cv2_ref.FB_PARAMS, threads=1)
            return torch.from_numpy(out.astype(np.float32))

        full = D.sharded_flow_series(compute_chunk, T, dst=0)
        if rank == 0:
            q.put(full.numpy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 10), (3, 10), (3, 3)])      # (3, 3): two pairs on three ranks, the last rank has no rows
def test_sharded_series_equals_unsharded(world, T):
    from btcs_pnes_optical_flow_b200 import synthetic as syn
    from oracle import cv2_ref
    spec = syn.ClipSpec(T=T, H=64, W=80, seed=1, patch=24, roi=32, amp=2.0)
    frames = syn.make_clip_np(spec)
    masks = np.stack([spec.roi_mask(), np.ones((64, 80), bool)])
    ref = cv2_ref.roi_series(frames, [1.0, 0.0], [0.0, 1.0], masks, cv2_ref.FB_PARAMS, threads=1).astype(np.float32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, masks, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == ref.shape and np.isnan(got[:, 0]).all()
    assert np.array_equal(got, ref, equal_nan=True)
