#!/usr/bin/env python
"""Golden vectors for the image pre/post-processing oracle (oracle/imageops_np.py) from the LIBRARIES the reference calls -- PIL
(api/app.py:143-150,199-203) and OpenCV (src/optimized_dataset.py:104-123) -- as installed in the build container.

    python tests/golden/make_imageops_golden.py      # writes tests/golden/imageops.npz (inputs + library outputs, ~60 KB)
"""
import os

import cv2
import numpy as np
import PIL
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
rs = np.random.RandomState(1234)


def smooth(h, w, c):
    """low-frequency pattern + noise, so that resampling differences are not hidden by pure noise"""
    y, x = np.mgrid[0:h, 0:w]
    base = 127 + 100 * np.sin(x / 7.0)[..., None] * np.cos(y / 5.0)[..., None] * np.ones(c)
    return np.clip(base + rs.randint(-25, 26, (h, w, c)), 0, 255).astype(np.uint8)


out = {"versions": np.array([f"PIL {PIL.__version__}", f"cv2 {cv2.__version__}"])}
rgb = smooth(45, 70, 3)
out["pil_rgb"] = rgb
out["pil_l"] = np.array(Image.fromarray(rgb).convert("L"))
for name, (ow, oh) in {"down": (32, 32), "up": (96, 64), "wonly": (35, 45), "honly": (70, 20)}.items():
    out[f"pil_{name}"] = np.array(Image.fromarray(rgb).convert("L").resize((ow, oh), Image.LANCZOS))
rgba = np.concatenate([rgb, rs.randint(0, 256, (45, 70, 1)).astype(np.uint8)], -1)
out["pil_rgba_l"] = np.array(Image.fromarray(rgba).convert("L"))
out["cv2_gray"] = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
g = out["cv2_gray"]
out["cv2_down"] = cv2.resize(g, (32, 32))
out["cv2_up"] = cv2.resize(g, (96, 64))
sq = smooth(64, 64, 1)[..., 0]
out["cv2_sq"] = sq
out["cv2_half"] = cv2.resize(sq, (32, 32))          # exact 2x: the INTER_AREA fast path
trip = smooth(40, 94, 3)                            # width not a multiple of 3
out["trip"] = trip
third = trip.shape[1] // 3
out["trip_gt"] = cv2.resize(cv2.cvtColor(np.ascontiguousarray(trip[:, :third]), cv2.COLOR_RGB2GRAY), (32, 32))
out["trip_glared"] = cv2.resize(cv2.cvtColor(np.ascontiguousarray(trip[:, third:2 * third]), cv2.COLOR_RGB2GRAY), (32, 32))
np.savez_compressed(os.path.join(HERE, "imageops.npz"), **out)
print("wrote imageops.npz", {k: v.shape for k, v in out.items()})
