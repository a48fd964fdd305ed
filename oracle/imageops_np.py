"""CPU oracle (TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this) for the image
pre/post-processing either side of the network:

  f1  /infer service path, api/app.py:136-157,190-203:  PIL `convert('L')`, PIL `resize(..., Image.LANCZOS)`, /255; clip*255 -> uint8,
      resize back.  PIL (pillow 11.2.1 pinned in requirements.txt:11) is not vendored in /root/reference, so the functions below restate
      Pillow's published algorithms (src/libImaging/Convert.c `rgb2l`, src/libImaging/Resample.c `precompute_coeffs`,
      `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`).  PINNED: tests/test_oracle.py checks them
      bit-for-bit against the installed PIL on random and real images, and against tests/golden/imageops.npz (made by
      tests/golden/make_imageops_golden.py with the PIL / cv2 of the build container).
  f2  training input pipeline, src/optimized_dataset.py:56-82,104-127,159-172:  triptych split (width // 3), cv2.cvtColor RGB2GRAY,
      cv2.resize (INTER_LINEAR on uint8), /255, then albumentations HorizontalFlip / RandomBrightnessContrast / GaussNoise.
      opencv (4.11 pinned, requirements.txt:9) is a third-party dependency as well: restated from modules/imgproc/src/color_rgb.simd.hpp
      (RGB2Gray<uchar>, 15-bit fixed point) and modules/imgproc/src/resize.cpp (HResizeLinear / VResizeLinear<uchar, int, short>, 11-bit
      coefficients; exact 2x down-scaling is INTER_AREA), PINNED the same way against the installed cv2.
      albumentations 2.0.6 is absent from this image: the augmentation FUNCTIONS (flip, alpha * x + beta clipped to [0,1], additive
      Gaussian noise clipped to [0,1]) are restated from its documentation; its random stream is not reproduced -- parity unpinned for
      the sampled parameters, exact for the functions given the parameters.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2   # Resample.c


def pil_rgb_to_l(rgb):
    """Convert.c rgb2l: L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16 (ITU-R 601-2 luma, 16-bit fixed point).  rgb [..., >=3] uint8."""
    r = rgb[..., 0].astype(np.int64)
    g = rgb[..., 1].astype(np.int64)
    b = rgb[..., 2].astype(np.int64)
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def pil_resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter (support 3) over the whole axis.
    Returns (bounds int32 [out, 2] = (first source index, tap count), kk int32 [out, ksize], ksize)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_axis(img, bounds, kk, axis):
    src = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(src.shape[:-1] + (bounds.shape[0],), np.uint8)
    for xx in range(bounds.shape[0]):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (src[..., x0:x0 + n] * kk[xx, :n].astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def pil_resize_lanczos(img, out_w, out_h):
    """Image.resize((out_w, out_h), Image.LANCZOS) of an 'L' image [H, W] uint8: horizontal pass (only the rows the vertical pass
    reads), then vertical pass, each rounding to uint8 (Resample.c ImagingResampleInner)."""
    in_h, in_w = img.shape
    need_h, need_v = out_w != in_w, out_h != in_h
    bh, kh, _ = pil_resample_coeffs(in_w, out_w)
    bv, kv, _ = pil_resample_coeffs(in_h, out_h)
    out = img
    if need_h:
        first = int(bv[0, 0])
        last = int(bv[-1, 0] + bv[-1, 1])
        out = _resample_axis(out[first:last], bh, kh, 1)
        bv = bv.copy()
        bv[:, 0] -= first
    if need_v:
        out = _resample_axis(out, bv, kv, 0)
    return np.ascontiguousarray(out)


# ---- OpenCV ------------------------------------------------------------------------------------------------------------------------
def cv2_rgb_to_gray(rgb):
    """cv2.cvtColor(rgb, COLOR_RGB2GRAY) on uint8: (R*9798 + G*19235 + B*3735 + 2^14) >> 15 (the library's 15-bit constants)."""
    r = rgb[..., 0].astype(np.int64)
    g = rgb[..., 1].astype(np.int64)
    b = rgb[..., 2].astype(np.int64)
    return ((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def cv2_linear_coeffs(in_size, out_size, clamp=True):
    """resize.cpp: source index and the two 11-bit taps per destination index (float32 arithmetic as in the library).  Along x the
    library clamps index AND weight at the borders; along y (clamp=False) it keeps the weights and clips the two ROW indices instead,
    so a border row is blended with itself through two separately truncated products."""
    scale = in_size / out_size
    ofs = np.zeros(out_size, np.int32)
    ab = np.zeros((out_size, 2), np.int32)
    for d in range(out_size):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(math.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        if clamp and sx < 0:
            fx, sx = np.float32(0), 0
        if clamp and sx >= in_size - 1:
            fx, sx = np.float32(0), in_size - 1
        ofs[d] = sx
        ab[d, 0] = int(np.rint(np.float32(np.float32(1.0) - fx) * np.float32(2048)))
        ab[d, 1] = int(np.rint(fx * np.float32(2048)))
    return ofs, ab


def cv2_resize_linear(img, out_w, out_h):
    """cv2.resize(img, (out_w, out_h)) (INTER_LINEAR) of a uint8 [H, W] image."""
    in_h, in_w = img.shape
    if in_w == out_w and in_h == out_h:
        return img.copy()
    if in_w == 2 * out_w and in_h == 2 * out_h:   # resize.cpp: exact 2x down-scaling runs the fast INTER_AREA path
        s = img.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xo, xa = cv2_linear_coeffs(in_w, out_w)
    yo, ya = cv2_linear_coeffs(in_h, out_h, clamp=False)
    s = img.astype(np.int32)
    x1 = np.minimum(xo + 1, in_w - 1)
    rows = s[:, xo] * xa[:, 0][None, :] + s[:, x1] * xa[:, 1][None, :]          # HResizeLinear: int, scale 2^11
    s0, s1 = rows[np.clip(yo, 0, in_h - 1)], rows[np.clip(yo + 1, 0, in_h - 1)]
    b0, b1 = ya[:, 0][:, None], ya[:, 1][:, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2        # VResizeLinear<uchar, int, short>
    return np.clip(out, 0, 255).astype(np.uint8)


def triptych_split_gray_resize(img_rgb, size):
    """optimized_dataset.py:104-123 for one RGB triptych [H, 3w(+r), 3] uint8 -> (glared, ground truth) uint8 [size, size]."""
    third = img_rgb.shape[1] // 3
    gt = cv2_rgb_to_gray(img_rgb[:, :third])
    gl = cv2_rgb_to_gray(img_rgb[:, third:2 * third])
    return cv2_resize_linear(gl, size, size), cv2_resize_linear(gt, size, size)


def augment(image, mask, flip, alpha, beta, noise):
    """optimized_dataset.py:159-172 given the sampled parameters: HorizontalFlip of image and mask; then on the image only
    RandomBrightnessContrast (clip(alpha * x + beta, 0, 1), brightness_by_max with max = 1 for float images) or additive noise
    (clip(x + noise, 0, 1)); float32 [H, W] in [0, 1].  `noise` is the already-sampled noise field or None."""
    image = image.astype(np.float32)
    mask = mask.astype(np.float32)
    if flip:
        image, mask = image[:, ::-1], mask[:, ::-1]
    if alpha != 1.0 or beta != 0.0:
        image = np.clip(image * np.float32(alpha) + np.float32(beta), 0.0, 1.0).astype(np.float32)
    if noise is not None:
        image = np.clip(image + noise.astype(np.float32), 0.0, 1.0).astype(np.float32)
    return np.ascontiguousarray(image), np.ascontiguousarray(mask)
