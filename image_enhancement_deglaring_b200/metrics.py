"""On-device validation metrics: PSNR and SSIM per image without leaving the GPU.

Mirrors what the reference computes on the host, image by image, with scikit-image:
  * `calculate_metrics(outputs, targets)` -- optimized_train.py:92-122 (first min(4, N) images, averages returned);
  * `psnr_ssim(outputs, targets, clip=True)` -- the per-image values evaluate.py:254-272 accumulates (output clipped to [0, 1]).
Both call `dg_image_metrics` (include/deglare.h); the reference's skimage defaults are fixed there (7x7 uniform window, K1 0.01,
K2 0.03, sample covariance, border cropped, data_range 1.0).
"""
import ctypes as C

import torch

from . import _lib


def psnr_ssim(outputs, targets, clip=False, data_range=1.0):
    """outputs, targets: float32 CUDA tensors [N,1,H,W] (or [N,H,W]).  Returns (psnr [N], ssim [N]) float64 CUDA tensors."""
    if not (outputs.is_cuda and targets.is_cuda):
        raise RuntimeError("psnr_ssim runs on CUDA tensors only; there is no CPU fallback")
    if outputs.shape != targets.shape:
        raise RuntimeError(f"shape mismatch: {tuple(outputs.shape)} vs {tuple(targets.shape)}")
    if outputs.dim() == 4:
        if outputs.shape[1] != 1:
            raise RuntimeError("metrics are defined for single-channel images")
        outputs, targets = outputs[:, 0], targets[:, 0]
    o = outputs.detach().float().contiguous()
    t = targets.detach().float().contiguous()
    N, H, W = o.shape
    acc = torch.zeros((N, 2), dtype=torch.float64, device=o.device)
    _lib.check(_lib.load().dg_image_metrics(o.data_ptr(), t.data_ptr(), N, H, W, 1 if clip else 0, float(data_range), acc.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
    psnr = 10.0 * torch.log10((data_range ** 2) * (H * W) / acc[:, 0])
    ssim = acc[:, 1] / float((H - 6) * (W - 6))
    return psnr, ssim


def calculate_metrics(outputs, targets):
    """optimized_train.py:92-122: mean PSNR and mean SSIM over the first min(4, N) images of the batch (two Python floats)."""
    n = min(4, outputs.size(0))
    psnr, ssim = psnr_ssim(outputs[:n], targets[:n])
    return float(psnr.mean()), float(ssim.mean())
