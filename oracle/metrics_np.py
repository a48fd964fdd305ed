"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the image-quality metrics the reference's validation / evaluation loops
compute per image (optimized_train.py:92-122, evaluate.py:254-272):

    skimage.metrics.peak_signal_noise_ratio(target, output, data_range=1.0)
    skimage.metrics.structural_similarity(target, output, data_range=1.0)

scikit-image (pinned 0.25.2 in the reference's requirements.txt:15) is NOT installed in this image and cannot be fetched, so
**parity for this row is unpinned**: what follows restates the published algorithm of those two functions for 2-D float32
inputs with default arguments -- win_size 7, uniform window (scipy.ndimage.uniform_filter, the same routine skimage calls),
K1 0.01, K2 0.03, sample covariance (NP/(NP-1)), float32 working type, the (win_size-1)//2 border cropped before the mean,
mean taken in float64 -- and `ssim_bruteforce` is an independent direct-window check of the same definition.
Only tests/ may import this module.
"""
import numpy as np
from scipy.ndimage import uniform_filter


def psnr(target, output, data_range=1.0):
    """peak_signal_noise_ratio: 10 log10(R^2 / mse), mse in float64."""
    err = np.mean((target.astype(np.float64) - output.astype(np.float64)) ** 2, dtype=np.float64)
    return 10.0 * np.log10((data_range ** 2) / err)


def ssim(target, output, data_range=1.0, win_size=7, K1=0.01, K2=0.03):
    """structural_similarity for 2-D float32 images, default arguments."""
    im1 = target.astype(np.float32)
    im2 = output.astype(np.float32)
    NP = win_size ** 2
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(im1, size=win_size)
    uy = uniform_filter(im2, size=win_size)
    uxx = uniform_filter(im1 * im1, size=win_size)
    uyy = uniform_filter(im2 * im2, size=win_size)
    uxy = uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (K1 * data_range) ** 2
    C2 = (K2 * data_range) ** 2
    A1, A2, B1, B2 = 2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))


def ssim_bruteforce(target, output, data_range=1.0, win_size=7, K1=0.01, K2=0.03):
    """The same definition evaluated window by window in float64 (small images only)."""
    x = target.astype(np.float64)
    y = output.astype(np.float64)
    H, W = x.shape
    NP = win_size ** 2
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    tot = 0.0
    for i in range(H - win_size + 1):
        for j in range(W - win_size + 1):
            a = x[i:i + win_size, j:j + win_size]
            b = y[i:i + win_size, j:j + win_size]
            ux, uy = a.mean(), b.mean()
            vx = ((a * a).mean() - ux * ux) * NP / (NP - 1)
            vy = ((b * b).mean() - uy * uy) * NP / (NP - 1)
            vxy = ((a * b).mean() - ux * uy) * NP / (NP - 1)
            tot += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return tot / ((H - win_size + 1) * (W - win_size + 1))
