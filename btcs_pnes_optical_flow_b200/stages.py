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


def level_image(plan: FlowPlan, frame: torch.Tensor, scale_index: int) -> torch.Tensor:
    """Pyramid level `scale_index` (0 = coarsest) of one full-resolution frame (uint8 or float32 [H, W])."""
    lib = _lib.load()
    if frame.dtype == torch.uint8:
        dtype = BF_DTYPE_U8
    else:
        frame, dtype = frame.float(), BF_DTYPE_F32
    frame = frame.contiguous()
    sc = plan.scales()[scale_index]
    out = torch.empty((sc["h"], sc["w"]), dtype=torch.float32, device=frame.device)
    check(lib.bf_stage_level_image(plan._h, frame.data_ptr(), dtype, frame.stride(0) * frame.element_size(),
                                   scale_index, out.data_ptr(), _stream(frame)))
    return out


def poly_exp(image: torch.Tensor, poly_n: int, poly_sigma: float) -> torch.Tensor:
    lib = _lib.load()
    image = image.float().contiguous()
    h, w = image.shape
    out = torch.empty((5, h, w), dtype=torch.float32, device=image.device)
    check(lib.bf_stage_poly_exp(image.data_ptr(), w, h, int(poly_n), float(poly_sigma), out.data_ptr(), _stream(image)))
    return out


def update_matrices(R0: torch.Tensor, R1: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    R0, R1, flow = R0.float().contiguous(), R1.float().contiguous(), flow.float().contiguous()
    _, h, w = R0.shape
    out = torch.empty((5, h, w), dtype=torch.float32, device=R0.device)
    check(lib.bf_stage_update_matrices(R0.data_ptr(), R1.data_ptr(), flow.data_ptr(), w, h, out.data_ptr(), _stream(R0)))
    return out


def blur_solve(M: torch.Tensor, winsize: int, flags: int = 0) -> torch.Tensor:
    lib = _lib.load()
    M = M.float().contiguous()
    _, h, w = M.shape
    out = torch.empty((h, w, 2), dtype=torch.float32, device=M.device)
    check(lib.bf_stage_blur_solve(M.data_ptr(), w, h, int(winsize), int(flags), out.data_ptr(), _stream(M)))
    return out


def upsample_flow(flow: torch.Tensor, w: int, h: int, mult: float) -> torch.Tensor:
    lib = _lib.load()
    flow = flow.float().contiguous()
    hs, ws, _ = flow.shape
    out = torch.empty((h, w, 2), dtype=torch.float32, device=flow.device)
    check(lib.bf_stage_upsample_flow(flow.data_ptr(), ws, hs, int(w), int(h), float(mult), out.data_ptr(), _stream(flow)))
    return out
