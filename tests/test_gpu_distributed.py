"""N > 1 on real GPUs: world_size-2 NCCL run of the temporal sharding with the CUDA flow series on each rank, against the
unsharded call on one GPU (chunk-boundary pair, ring-slot reuse across chunks, gather over NVLink).  Needs 2 devices."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import btcs_pnes_optical_flow_b200 as B
        from btcs_pnes_optical_flow_b200 import distributed as D, synthetic as syn
        spec = syn.ClipSpec(T=T, H=270, W=480, seed=5, patch=90, roi=120, amp=4.0)
        masks = torch.from_numpy(np.stack([spec.roi_mask(), np.ones((270, 480), bool)])).cuda()
        plan = B.FlowPlan(480, 270, B.FB_PARAMS, max_pairs=4, max_rois=2, device=rank)     # chunks longer than a batch: ring reuse

        def compute_chunk(f0, f1):
            return plan.flow_series(syn.make_clip(spec, f"cuda:{rank}", f0, f1 - f0), None, None, masks)

        full = D.sharded_flow_series(compute_chunk, T, dst=0)
        if rank == 0:
            whole = plan.flow_series(syn.make_clip(spec, "cuda:0", 0, T), None, None, masks)
            q.put((full.cpu().numpy(), whole.cpu().numpy()))
        else:
            assert full is None
        plan.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("T", [23, 3])
def test_nccl_sharded_series_equals_unsharded(T):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got.shape == want.shape == (2, T, 3) and np.isnan(got[:, 0]).all()
    # each pair is computed by the same kernels on the same frames whichever rank owns it: bit-identical rows
    assert np.array_equal(got, want, equal_nan=True)
