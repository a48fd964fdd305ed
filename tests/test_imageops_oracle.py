"""CPU: the image pre/post-processing oracle (oracle/imageops_np.py) is pinned against the golden vectors from PIL / OpenCV
(tests/golden/imageops.npz, made by tests/golden/make_imageops_golden.py) and, where those libraries are importable, against the
libraries themselves; the product's host-side tap tables equal the oracle's."""
import os

import numpy as np
import pytest

from oracle import imageops_np as io

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "imageops.npz"))


def test_pil_path_matches_golden(gold):
    rgb = gold["pil_rgb"]
    l = io.pil_rgb_to_l(rgb)
    assert np.array_equal(l, gold["pil_l"])
    rgba = np.concatenate([rgb, np.zeros_like(rgb[..., :1])], -1)
    assert np.array_equal(io.pil_rgb_to_l(rgba), gold["pil_rgba_l"])          # alpha is ignored
    for name, (ow, oh) in {"down": (32, 32), "up": (96, 64), "wonly": (35, 45), "honly": (70, 20)}.items():
        assert np.array_equal(io.pil_resize_lanczos(l, ow, oh), gold[f"pil_{name}"]), name


def test_cv2_path_matches_golden(gold):
    g = io.cv2_rgb_to_gray(gold["pil_rgb"])
    assert np.array_equal(g, gold["cv2_gray"])
    assert np.array_equal(io.cv2_resize_linear(g, 32, 32), gold["cv2_down"])
    assert np.array_equal(io.cv2_resize_linear(g, 96, 64), gold["cv2_up"])
    assert np.array_equal(io.cv2_resize_linear(gold["cv2_sq"], 32, 32), gold["cv2_half"])   # exact 2x = INTER_AREA
    gl, gt = io.triptych_split_gray_resize(gold["trip"], 32)
    assert np.array_equal(gl, gold["trip_glared"]) and np.array_equal(gt, gold["trip_gt"])


def test_oracle_matches_installed_pil():
    Image = pytest.importorskip("PIL.Image")
    rs = np.random.RandomState(0)
    for (h, w, oh, ow) in [(37, 53, 64, 64), (300, 200, 128, 128), (700, 500, 96, 160), (64, 64, 200, 120)]:
        rgb = rs.randint(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(rgb).convert("L").resize((ow, oh), Image.LANCZOS))
        assert np.array_equal(io.pil_resize_lanczos(io.pil_rgb_to_l(rgb), ow, oh), ref), (h, w, oh, ow)


def test_oracle_matches_installed_cv2():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(1)
    for (h, w, oh, ow) in [(37, 53, 64, 64), (300, 200, 128, 128), (256, 256, 128, 128), (333, 517, 96, 96), (5, 7, 16, 16)]:
        rgb = rs.randint(0, 256, (h, w, 3), dtype=np.uint8)
        g = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
        assert np.array_equal(io.cv2_rgb_to_gray(rgb), g)
        assert np.array_equal(io.cv2_resize_linear(g, ow, oh), cv2.resize(g, (ow, oh))), (h, w, oh, ow)


def test_product_tap_tables_equal_the_oracle():
    from image_enhancement_deglaring_b200 import imageops as prod
    for (i, o) in [(53, 64), (700, 96), (4096, 512), (512, 333), (64, 64)]:
        b, k, ks = prod._pil_lanczos_host(i, o)
        ob, ok, oks = io.pil_resample_coeffs(i, o)
        assert ks == oks and np.array_equal(b, ob) and np.array_equal(k, ok), (i, o)
        assert (k.sum(1) - (1 << 22)).__abs__().max() <= k.shape[1]                # rows sum to 1 up to rounding
        for clamp in (True, False):
            xo, xa = prod._cv2_linear_host(i, o, clamp)
            oo, oa = io.cv2_linear_coeffs(i, o, clamp)
            assert np.array_equal(xo, oo) and np.array_equal(xa, oa), (i, o, clamp)


def test_augment_oracle_properties():
    rs = np.random.RandomState(2)
    img = rs.rand(8, 12).astype(np.float32)
    msk = rs.rand(8, 12).astype(np.float32)
    a, m = io.augment(img, msk, True, 1.0, 0.0, None)
    assert np.array_equal(a, img[:, ::-1]) and np.array_equal(m, msk[:, ::-1])
    a, m = io.augment(img, msk, False, 1.2, -0.1, None)
    assert np.array_equal(m, msk) and a.min() >= 0 and a.max() <= 1
    assert np.allclose(a, np.clip(img * 1.2 - 0.1, 0, 1), atol=1e-6)
