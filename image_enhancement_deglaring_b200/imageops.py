"""Image pre/post-processing either side of the network on the device, bit-exact with the host libraries the
reference calls:

  f1  /infer (api/app.py:136-157,190-203): `infer_preprocess` = PIL convert('L') + resize((512, 512), LANCZOS) -> uint8 [N,1,512,512]
      (feeds `LightweightUNet.forward_u8`, whose first kernel does the /255 and whose head does clip*255 -> uint8), `resize_back` =
      the LANCZOS resize of the uint8 result to the upload's size; `infer_image` chains the whole request.
  f2  training input (src/optimized_dataset.py:104-127,159-172): `triptych_to_pairs` = split at width // 3, cv2 RGB2GRAY,
      cv2.resize INTER_LINEAR; `sample_augment_params` + `augment` = HorizontalFlip / RandomBrightnessContrast / GaussNoise.

Only the tap tables are built on the host (once per size pair, in the libraries' own double / float arithmetic); every pixel is
touched by CUDA kernels (csrc/imageops.cu) through the C-ABI.  There is no CPU fallback.
"""
import functools
import math

import numpy as np
import torch

from . import _lib

_PIL_BITS = 22


def _sinc(x):
    if x == 0.0:
        return 1.0
    x *= math.pi
    return math.sin(x) / x


@functools.lru_cache(maxsize=64)
def _pil_lanczos_host(in_size, out_size):
    """Pillow Resample.c precompute_coeffs (LANCZOS, support 3) + normalize_coeffs_8bpc: bounds [out, 2], kk [out, ksize] int32."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 3.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / fscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    one = float(1 << _PIL_BITS)
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        taps = []
        total = 0.0
        for i in range(lo, hi):
            t = (i - center + 0.5) * inv
            w = _sinc(t) * _sinc(t / 3) if -3.0 <= t < 3.0 else 0.0
            taps.append(w)
            total += w
        for j, w in enumerate(taps):
            if total != 0.0:
                w = w / total
            kk[o, j] = int(w * one - 0.5) if w < 0 else int(w * one + 0.5)
        bounds[o] = (lo, hi - lo)
    return bounds, kk, ksize


@functools.lru_cache(maxsize=64)
def _cv2_linear_host(in_size, out_size, clamp):
    """OpenCV resize.cpp INTER_LINEAR tables for uint8: source index and the two taps scaled by 2^11 (float32 arithmetic, round half
    to even).  Along x (clamp=True) index and weight are clamped at the borders; along y the kernel clips the row indices instead."""
    scale = in_size / out_size
    ofs = np.zeros(out_size, np.int32)
    ab = np.zeros((out_size, 2), np.int32)
    f32 = np.float32
    for d in range(out_size):
        f = f32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = f32(f - f32(s))
        if clamp and s < 0:
            f, s = f32(0), 0
        if clamp and s >= in_size - 1:
            f, s = f32(0), in_size - 1
        ofs[d] = s
        ab[d] = (int(np.rint(f32(f32(1) - f) * f32(2048))), int(np.rint(f * f32(2048))))
    return ofs, ab


_dev_tables = {}


def _on_device(key, build, device):
    k = (key, str(device))
    if k not in _dev_tables:
        _dev_tables[k] = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in build())
    return _dev_tables[k]


def pil_lanczos_tables(in_size, out_size, device):
    """(bounds, kk) int32 device tensors + ksize for one axis (cached per size pair and device)."""
    b, k, ksize = _pil_lanczos_host(in_size, out_size)
    tb, tk = _on_device(("pil", in_size, out_size), lambda: (b, k), device)
    return tb, tk, ksize


def cv2_linear_tables(in_size, out_size, device, clamp):
    return _on_device(("cv2", in_size, out_size, clamp), lambda: _cv2_linear_host(in_size, out_size, clamp), device)


def _as_batch_u8(img):
    if not (isinstance(img, torch.Tensor) and img.is_cuda and img.dtype == torch.uint8):
        raise RuntimeError("imageops: a CUDA uint8 tensor is required (there is no CPU fallback)")
    if img.dim() == 2:
        img = img[None, :, :, None]
    elif img.dim() == 3:
        img = img[None] if img.shape[-1] in (1, 3, 4) else img[..., None]
    if img.dim() != 4 or img.shape[-1] not in (1, 3, 4):
        raise RuntimeError(f"imageops: expected [H,W], [H,W,C] or [N,H,W,C] with C in (1,3,4), got {tuple(img.shape)}")
    return img.contiguous()


def pil_resize(img, out_w, out_h):
    """PIL `Image.fromarray(img).convert('L').resize((out_w, out_h), Image.LANCZOS)` for uint8 CUDA images [H,W], [H,W,C] or [N,H,W,C]
    (C = 1, 3 (RGB) or 4 (RGBA)); returns uint8 [N, out_h, out_w].  api/app.py:143-150, :199-203."""
    img = _as_batch_u8(img)
    N, in_h, in_w, C = img.shape
    dev = img.device
    with torch.cuda.device(dev):
        out = torch.empty((N, out_h, out_w), dtype=torch.uint8, device=dev)
        bh = kh = bv = kv = None
        ksh = ksv = 0
        row0, rows = 0, in_h
        if out_w != in_w:
            bh, kh, ksh = pil_lanczos_tables(in_w, out_w, dev)
        tmp = None
        if out_h != in_h:
            bv, kv, ksv = pil_lanczos_tables(in_h, out_h, dev)
            hb = _pil_lanczos_host(in_h, out_h)[0]
            row0 = int(hb[0, 0])
            rows = int(hb[-1, 0] + hb[-1, 1]) - row0
            tmp = torch.empty((N, rows, out_w), dtype=torch.uint8, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        _lib.check(_lib.load().dg_pil_resize_u8(img.data_ptr(), C, N, in_h, in_w, out.data_ptr(), out_h, out_w, ptr(bh), ptr(kh), ksh,
                                               ptr(bv), ptr(kv), ksv, row0, rows, ptr(tmp), 0 if tmp is None else tmp.numel(),
                                               torch.cuda.current_stream().cuda_stream))
    return out


def infer_preprocess(img, size=512):
    """The pre-processing of one /infer request (api/app.py:136-157) up to the uint8 network input [N,1,size,size]."""
    return pil_resize(img, size, size)[:, None]


def resize_back(y_u8, original_size):
    """api/app.py:199-203: the uint8 result [N,1,H,W] (or [N,H,W]) resized to original_size = (width, height) with LANCZOS."""
    if y_u8.dim() == 4:
        y_u8 = y_u8[:, 0]
    w, h = original_size
    return pil_resize(y_u8[..., None], int(w), int(h))


def infer_image(net, img, size=512):
    """Everything /infer computes between PNG decode and PNG encode (api/app.py:136-203) on the device: gray, LANCZOS resize, /255,
    the network, clip*255 -> uint8, LANCZOS resize back to the upload's size.  img: CUDA uint8 [H,W] / [H,W,C] / [N,H,W,C]."""
    b = _as_batch_u8(img)
    x = infer_preprocess(b, size)
    with torch.no_grad():
        y = net.forward_u8(x)
    return resize_back(y, (b.shape[2], b.shape[1]))


def cv2_resize(img, out_w, out_h, x_off=0, width=None):
    """cv2.cvtColor(COLOR_RGB2GRAY) (if C == 3) + cv2.resize(..., (out_w, out_h)) of the column panel [x_off, x_off + width) of uint8
    CUDA images [N,H,W,C] / [H,W,C] / [H,W]; returns uint8 [N, out_h, out_w].  src/optimized_dataset.py:113-123."""
    img = _as_batch_u8(img)
    N, in_h, in_wf, C = img.shape
    if C == 4:
        raise RuntimeError("cv2_resize: 1 or 3 channels")
    width = in_wf - x_off if width is None else width
    dev = img.device
    with torch.cuda.device(dev):
        out = torch.empty((N, out_h, out_w), dtype=torch.uint8, device=dev)
        if width == out_w and in_h == out_h and C == 1:
            return img[:, :, x_off:x_off + width, 0].contiguous()
        xo, xa = cv2_linear_tables(width, out_w, dev, True)
        yo, ya = cv2_linear_tables(in_h, out_h, dev, False)
        _lib.check(_lib.load().dg_cv2_resize_u8(img.data_ptr(), C, N, in_h, in_wf, x_off, width, out.data_ptr(), out_h, out_w,
                                               xo.data_ptr(), xa.data_ptr(), yo.data_ptr(), ya.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
    return out


def triptych_to_pairs(img_rgb, size=512):
    """src/optimized_dataset.py:104-123: RGB triptychs [N,H,3w(+r),3] uint8 (ground truth | glared | ...) -> (glared, ground truth)
    uint8 [N,size,size] each: split at width // 3, gray, resize."""
    img = _as_batch_u8(img_rgb)
    third = img.shape[2] // 3
    return cv2_resize(img, size, size, third, third), cv2_resize(img, size, size, 0, third)


def sample_augment_params(n, generator=None):
    """Parameters of get_optimized_transformations' training pipeline (src/optimized_dataset.py:159-172, albumentations 2.0.6
    defaults) for n samples, float32 [n,4] = (flip, alpha, beta, sigma): HorizontalFlip p = .5; OneOf p = .5 of
    RandomBrightnessContrast (weight .8; alpha = 1 + U(-.2,.2), beta = U(-.2,.2)) and GaussNoise (weight .2; sigma = U(.2,.44)).
    The random stream is torch's, not albumentations'."""
    u = torch.rand((n, 6), generator=generator)
    p = torch.zeros((n, 4), dtype=torch.float32)
    p[:, 0] = (u[:, 0] < 0.5).float()
    p[:, 1] = 1.0
    one_of = u[:, 1] < 0.5
    bc = one_of & (u[:, 2] < 0.8)
    gn = one_of & ~(u[:, 2] < 0.8)
    p[bc, 1] = 1.0 + (u[bc, 3] * 0.4 - 0.2)
    p[bc, 2] = u[bc, 4] * 0.4 - 0.2
    p[gn, 3] = 0.2 + u[gn, 5] * 0.24
    return p


def augment(image_u8, mask_u8, params, seed=0):
    """uint8 [N,H,W] image / mask -> float32 [N,1,H,W] pair: /255, flip (both), brightness-contrast and noise (image only); params as
    from `sample_augment_params`.  src/optimized_dataset.py:126-141."""
    if not (image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.dim() == 3):
        raise RuntimeError("augment: CUDA uint8 [N,H,W] tensors are required")
    N, H, W = image_u8.shape
    dev = image_u8.device
    image_u8 = image_u8.contiguous()
    mask_u8 = None if mask_u8 is None else mask_u8.contiguous()
    with torch.cuda.device(dev):
        prm = params.to(device=dev, dtype=torch.float32).contiguous()
        out_i = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev)
        out_m = None if mask_u8 is None else torch.empty((N, 1, H, W), dtype=torch.float32, device=dev)
        _lib.check(_lib.load().dg_augment(image_u8.data_ptr(), None if mask_u8 is None else mask_u8.data_ptr(), out_i.data_ptr(),
                                         None if out_m is None else out_m.data_ptr(), N, H, W, prm.data_ptr(), int(seed) & (2 ** 64 - 1),
                                         torch.cuda.current_stream().cuda_stream))
    return out_i, out_m
