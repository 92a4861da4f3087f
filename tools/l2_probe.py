"""Effective L2 capacity probe: read-only and read-write sweeps over buffers of growing size (GB/s vs footprint)."""
import torch
torch.cuda.init()
dev = torch.device("cuda")
def bw(nbytes, rw, iters=60):
    n = nbytes // 4
    x = torch.ones(n, dtype=torch.float32, device=dev)
    y = torch.empty_like(x) if rw else None
    for _ in range(5):
        (torch.mul(x, 1.0001, out=y) if rw else x.sum())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        (torch.mul(x, 1.0001, out=y) if rw else x.sum())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return nbytes * (2 if rw else 1) / ms / 1e6, ms * 1e3
for mb in (8, 16, 24, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256, 512):
    r, tr = bw(mb << 20, False)
    w, tw = bw(mb << 20, True)
    print(f"footprint {mb:4d} MB  read-only {r:7.0f} GB/s ({tr:6.1f} us)   | x->y copy-scale footprint {2*mb:4d} MB {w:7.0f} GB/s ({tw:6.1f} us)")
