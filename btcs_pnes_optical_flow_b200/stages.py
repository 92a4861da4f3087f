"""Stage-level access to the kernels (bf_stage_* in include/btcsflow.h) on torch CUDA tensors.

No reference counterpart (cv2 exposes only the final flow); exists so each kernel can be compared with the
stage-level oracle.  Layouts follow the device layout: R and M are planes [5, h, w]; flow is [h, w, 2].
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import BF_DTYPE_F32, BF_DTYPE_U8, check
from .flow import FlowPlan


def _stream(t: torch.Tensor) -> int:
    return int(torch.cuda.current_stream(t.device).cuda_stream)


# Output buffers are carved out of a larger allocation with canary bands on both sides: compute-sanitizer is not
# available on every pool, so out-of-bounds WRITES of a kernel are caught by `check_guards()` in the tests instead.
_GUARD = 4096
_guards: list = []


def _out(shape, device) -> torch.Tensor:
    n = 1
    for d in shape:
        n *= int(d)
    buf = torch.full((n + 2 * _GUARD,), -12345.0, dtype=torch.float32, device=device)
    _guards.append((buf, n))
    if len(_guards) > 256:                 # bounded: callers that never check do not accumulate device memory
        _guards.pop(0)
    return buf[_GUARD:_GUARD + n].view(*shape)


def check_guards() -> int:
    """Assert no stage kernel wrote outside its output buffer; returns the number of buffers checked."""
    torch.cuda.synchronize()
    k = len(_guards)
    for buf, n in _guards:
        lo, hi = buf[:_GUARD], buf[_GUARD + n:]
        assert bool((lo == -12345.0).all()) and bool((hi == -12345.0).all()), "out-of-bounds write detected"
    _guards.clear()
    return k


def level_image(plan: FlowPlan, frame: torch.Tensor, scale_index: int) -> torch.Tensor:
    """Pyramid level `scale_index` (0 = coarsest) of one full-resolution frame (uint8 or float32 [H, W])."""
    lib = _lib.load()
    if frame.dtype == torch.uint8:
        dtype = BF_DTYPE_U8
    else:
        frame, dtype = frame.float(), BF_DTYPE_F32
    frame = frame.contiguous()
    sc = plan.scales()[scale_index]
    out = _out((sc["h"], sc["w"]), frame.device)
    check(lib.bf_stage_level_image(plan._h, frame.data_ptr(), dtype, frame.stride(0) * frame.element_size(),
                                   scale_index, out.data_ptr(), _stream(frame)))
    return out


def poly_exp(image: torch.Tensor, poly_n: int, poly_sigma: float) -> torch.Tensor:
    lib = _lib.load()
    image = image.float().contiguous()
    h, w = image.shape
    out = _out((5, h, w), image.device)
    check(lib.bf_stage_poly_exp(image.data_ptr(), w, h, int(poly_n), float(poly_sigma), out.data_ptr(), _stream(image)))
    return out


def update_matrices(R0: torch.Tensor, R1: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    R0, R1, flow = R0.float().contiguous(), R1.float().contiguous(), flow.float().contiguous()
    _, h, w = R0.shape
    out = _out((5, h, w), R0.device)
    check(lib.bf_stage_update_matrices(R0.data_ptr(), R1.data_ptr(), flow.data_ptr(), w, h, out.data_ptr(), _stream(R0)))
    return out


def blur_solve(M: torch.Tensor, winsize: int, flags: int = 0) -> torch.Tensor:
    lib = _lib.load()
    M = M.float().contiguous()
    _, h, w = M.shape
    out = _out((h, w, 2), M.device)
    check(lib.bf_stage_blur_solve(M.data_ptr(), w, h, int(winsize), int(flags), out.data_ptr(), _stream(M)))
    return out


def upsample_flow(flow: torch.Tensor, w: int, h: int, mult: float) -> torch.Tensor:
    lib = _lib.load()
    flow = flow.float().contiguous()
    hs, ws, _ = flow.shape
    out = _out((h, w, 2), flow.device)
    check(lib.bf_stage_upsample_flow(flow.data_ptr(), ws, hs, int(w), int(h), float(mult), out.data_ptr(), _stream(flow)))
    return out
