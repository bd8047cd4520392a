"""Device-side post-processing of SR frames: the reference converts every output frame to a
uint8 BGR image on the CPU (lbasicsr/utils/img_util.py:38-94) and computes PSNR and SSIM on the Y channel with numpy / cv2
(lbasicsr/metrics/psnr_ssim.py:11-48, 85-129, 172-200); here they run as CUDA kernels on the frames still resident in HBM."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _capi as K
from . import engine


def tensor2img_psnr(sr: torch.Tensor, gt: Optional[torch.Tensor] = None, want_image: bool = True
                    ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """sr, gt: float32 [n,3,H,W] RGB on a CUDA device.  Returns (uint8 [n,H,W,3] BGR images or None,
    float64 [n] PSNR-Y in dB or None).  Identical frames give inf, like the reference."""
    if not sr.is_cuda:
        raise RuntimeError("savsr_b200.postproc runs on CUDA only; there is no CPU fallback")
    if sr.dim() != 4 or sr.shape[1] != 3:
        raise ValueError(f"expected [n,3,H,W], got {tuple(sr.shape)}")
    if gt is not None and gt.shape != sr.shape:
        raise AssertionError(f"Image shapes are different: {tuple(sr.shape)}, {tuple(gt.shape)}.")   # psnr_ssim.py:26
    sr = sr.float().contiguous()
    gt = gt.to(sr.device).float().contiguous() if gt is not None else None
    n, _, H, W = sr.shape
    ctx = engine.context(sr.device.index if sr.device.index is not None else torch.cuda.current_device())
    img = torch.empty(n, H, W, 3, dtype=torch.uint8, device=sr.device) if want_image else None
    nb = ctx.lib.savsr_img_metrics_blocks(ctx.handle, H, W)
    part = torch.empty(n, nb, dtype=torch.float64, device=sr.device) if gt is not None else None
    with torch.cuda.device(sr.device):
        K.check(ctx.lib.savsr_img_metrics(ctx.handle, sr.data_ptr(), gt.data_ptr() if gt is not None else None, n, H, W,
                                          img.data_ptr() if img is not None else None, part.data_ptr() if part is not None else None,
                                          torch.cuda.current_stream().cuda_stream))
    psnr = None
    if part is not None:
        mse = part.sum(dim=1) / float(H * W)          # per-block partials, fixed summation order: bit-reproducible
        psnr = torch.where(mse == 0, torch.full_like(mse, math.inf), 10.0 * torch.log10(255.0 * 255.0 / mse))
    return img, psnr


def ssim_y(sr: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """calculate_ssim(crop_border=0, test_y_channel=True) of the reference (the YAML metric `ssim_y`) per frame:
    sr, gt float32 [n,3,H,W] RGB on a CUDA device -> float64 [n]."""
    if not sr.is_cuda:
        raise RuntimeError("savsr_b200.postproc runs on CUDA only; there is no CPU fallback")
    if sr.dim() != 4 or sr.shape[1] != 3:
        raise ValueError(f"expected [n,3,H,W], got {tuple(sr.shape)}")
    if gt.shape != sr.shape:
        raise AssertionError(f"Image shapes are different: {tuple(sr.shape)}, {tuple(gt.shape)}.")   # psnr_ssim.py:109
    sr = sr.float().contiguous()
    gt = gt.to(sr.device).float().contiguous()
    n, _, H, W = sr.shape
    ctx = engine.context(sr.device.index if sr.device.index is not None else torch.cuda.current_device())
    nb = ctx.lib.savsr_ssim_y_blocks(H, W)
    if nb == 0:
        raise ValueError(f"frames of {H}x{W} are smaller than the 11x11 SSIM window")
    part = torch.empty(n, nb, dtype=torch.float64, device=sr.device)
    with torch.cuda.device(sr.device):
        K.check(ctx.lib.savsr_ssim_y(ctx.handle, sr.data_ptr(), gt.data_ptr(), n, H, W, part.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream))
    return part.sum(dim=1) / float((H - 10) * (W - 10))
