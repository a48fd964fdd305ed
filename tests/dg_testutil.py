import numpy as np


def psnr(a, b, data_range=1.0):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(data_range ** 2 / mse)
