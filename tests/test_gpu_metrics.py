"""On-device PSNR / SSIM (dg_image_metrics) against the CPU restatement of the skimage calls the reference makes
(optimized_train.py:92-122, evaluate.py:254-272)."""
import numpy as np
import pytest
import torch

from oracle import metrics_np as M

pytestmark = pytest.mark.gpu

dg = pytest.importorskip("image_enhancement_deglaring_b200")


def _pair(shape, seed, noise=0.05):
    rs = np.random.RandomState(seed)
    t = rs.rand(*shape).astype(np.float32)
    # smooth the target a little so windows have structure, then perturb
    t = (t + np.roll(t, 1, -1) + np.roll(t, 1, -2) + np.roll(t, (1, 1), (-2, -1))) / 4
    o = (t + noise * rs.standard_normal(shape)).astype(np.float32)
    return o, t


@pytest.mark.parametrize("shape", [(3, 1, 64, 64), (2, 1, 48, 80), (1, 1, 512, 512), (2, 1, 7, 9), (1, 1, 23, 40)])
def test_psnr_ssim_match_oracle(shape):
    from image_enhancement_deglaring_b200.metrics import psnr_ssim
    o, t = _pair(shape, 9)
    for clip in (False, True):
        psnr, ssim = psnr_ssim(torch.from_numpy(o).cuda(), torch.from_numpy(t).cuda(), clip=clip)
        for i in range(shape[0]):
            oi = np.clip(o[i, 0], 0, 1) if clip else o[i, 0]
            assert abs(float(psnr[i]) - M.psnr(t[i, 0], oi)) <= 1e-4, (shape, clip, i)
            assert abs(float(ssim[i]) - M.ssim(t[i, 0], oi)) <= 1e-4, (shape, clip, i)


def test_calculate_metrics_mirrors_training_loop_helper():
    """optimized_train.py:92-122: first min(4, N) images, averaged."""
    from image_enhancement_deglaring_b200.metrics import calculate_metrics
    o, t = _pair((6, 1, 64, 64), 10)
    p, s = calculate_metrics(torch.from_numpy(o).cuda(), torch.from_numpy(t).cuda())
    want_p = np.mean([M.psnr(t[i, 0], o[i, 0]) for i in range(4)])
    want_s = np.mean([M.ssim(t[i, 0], o[i, 0]) for i in range(4)])
    assert abs(p - want_p) <= 1e-4 and abs(s - want_s) <= 1e-4


def test_metrics_errors():
    from image_enhancement_deglaring_b200.metrics import psnr_ssim
    with pytest.raises(RuntimeError):
        psnr_ssim(torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))          # CPU tensors
    with pytest.raises(RuntimeError):
        psnr_ssim(torch.zeros(1, 1, 4, 16).cuda(), torch.zeros(1, 1, 4, 16).cuda())  # smaller than the window
    with pytest.raises(RuntimeError):
        psnr_ssim(torch.zeros(1, 1, 16, 16).cuda(), torch.zeros(1, 1, 16, 8).cuda())
