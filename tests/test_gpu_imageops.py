"""Image pre/post-processing on the device (csrc/imageops.cu through the C-ABI) against the CPU oracle
(oracle/imageops_np.py, pinned to PIL / OpenCV) -- integer work, so the bar is bit-exact -- and, where importable, against the
libraries the reference calls themselves (api/app.py:143-150,199-203; src/optimized_dataset.py:104-127,159-172)."""
import os

import numpy as np
import pytest
import torch

from oracle import imageops_np as io

pytestmark = pytest.mark.gpu

dg = pytest.importorskip("image_enhancement_deglaring_b200")
HERE = os.path.dirname(os.path.abspath(__file__))


def _img(h, w, c, seed):
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = 127 + 100 * np.sin(x / 9.0)[..., None] * np.cos(y / 6.0)[..., None] * np.ones(c)
    return np.clip(base + rs.randint(-40, 41, (h, w, c)), 0, 255).astype(np.uint8)


PIL_CASES = [(45, 70, 3, 32, 32), (300, 200, 3, 512, 512), (1024, 768, 1, 512, 512), (512, 512, 1, 300, 200), (700, 512, 4, 512, 512),
             (512, 900, 3, 512, 512), (2048, 1536, 3, 512, 512), (16, 16, 1, 16, 16)]


@pytest.mark.parametrize("h,w,c,oh,ow", PIL_CASES)
def test_pil_gray_lanczos_resize_is_bit_exact(h, w, c, oh, ow):
    from image_enhancement_deglaring_b200 import imageops
    img = _img(h, w, c, h + w)
    got = imageops.pil_resize(torch.from_numpy(img if c > 1 else img[..., 0]).cuda(), ow, oh)[0].cpu().numpy()
    l = io.pil_rgb_to_l(img) if c > 1 else img[..., 0]
    assert np.array_equal(got, io.pil_resize_lanczos(l, ow, oh)), "device vs oracle"
    Image = pytest.importorskip("PIL.Image")
    src = Image.fromarray(img if c > 1 else img[..., 0])
    ref = np.array(src.convert("L").resize((ow, oh), Image.LANCZOS))
    assert np.array_equal(got, ref), "device vs PIL"


def test_pil_resize_golden_and_batch():
    from image_enhancement_deglaring_b200 import imageops
    g = np.load(os.path.join(HERE, "golden", "imageops.npz"))
    rgb = torch.from_numpy(g["pil_rgb"]).cuda()
    for name, (ow, oh) in {"down": (32, 32), "up": (96, 64), "wonly": (35, 45), "honly": (70, 20)}.items():
        assert np.array_equal(imageops.pil_resize(rgb, ow, oh)[0].cpu().numpy(), g[f"pil_{name}"]), name
    assert np.array_equal(imageops.pil_resize(rgb, 70, 45)[0].cpu().numpy(), g["pil_l"])       # convert('L') only
    batch = torch.stack([rgb, rgb.flip(0), rgb.flip(1)])
    out = imageops.pil_resize(batch, 32, 32).cpu().numpy()
    for i in range(3):
        assert np.array_equal(out[i], io.pil_resize_lanczos(io.pil_rgb_to_l(batch[i].cpu().numpy()), 32, 32))


CV2_CASES = [(45, 70, 3, 32, 32), (300, 200, 1, 512, 512), (1024, 1024, 3, 512, 512), (600, 800, 3, 512, 512), (333, 517, 1, 512, 512),
             (5, 7, 1, 16, 16), (512, 512, 1, 512, 512)]


@pytest.mark.parametrize("h,w,c,oh,ow", CV2_CASES)
def test_cv2_gray_linear_resize_is_bit_exact(h, w, c, oh, ow):
    from image_enhancement_deglaring_b200 import imageops
    img = _img(h, w, c, 3 * h + w)
    got = imageops.cv2_resize(torch.from_numpy(img if c > 1 else img[..., 0]).cuda(), ow, oh)[0].cpu().numpy()
    gray = io.cv2_rgb_to_gray(img) if c > 1 else img[..., 0]
    assert np.array_equal(got, io.cv2_resize_linear(gray, ow, oh)), "device vs oracle"
    cv2 = pytest.importorskip("cv2")
    ref = cv2.resize(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) if c > 1 else img[..., 0], (ow, oh))
    assert np.array_equal(got, ref), "device vs cv2"


def test_triptych_split_matches_the_dataset_code():
    """src/optimized_dataset.py:104-123 on a batch of triptychs whose width is not a multiple of three."""
    from image_enhancement_deglaring_b200 import imageops
    g = np.load(os.path.join(HERE, "golden", "imageops.npz"))
    gl, gt = imageops.triptych_to_pairs(torch.from_numpy(g["trip"]).cuda(), 32)
    assert np.array_equal(gl[0].cpu().numpy(), g["trip_glared"]) and np.array_equal(gt[0].cpu().numpy(), g["trip_gt"])
    trips = np.stack([_img(200, 3 * 170 + 2, 3, s) for s in range(3)])
    gl, gt = imageops.triptych_to_pairs(torch.from_numpy(trips).cuda(), 128)
    for i in range(3):
        ogl, ogt = io.triptych_split_gray_resize(trips[i], 128)
        assert np.array_equal(gl[i].cpu().numpy(), ogl) and np.array_equal(gt[i].cpu().numpy(), ogt)


def test_augment_functions_match_the_oracle_and_noise_has_the_asked_statistics():
    from image_enhancement_deglaring_b200 import imageops
    rs = np.random.RandomState(5)
    img = rs.randint(0, 256, (4, 64, 96), dtype=np.uint8)
    msk = rs.randint(0, 256, (4, 64, 96), dtype=np.uint8)
    params = torch.tensor([[0, 1.0, 0.0, 0.0], [1, 1.0, 0.0, 0.0], [0, 1.17, -0.08, 0.0], [1, 0.83, 0.15, 0.0]], dtype=torch.float32)
    oi, om = imageops.augment(torch.from_numpy(img).cuda(), torch.from_numpy(msk).cuda(), params)
    for n in range(4):
        f, a, b, _ = params[n].tolist()
        ri, rm = io.augment(img[n].astype(np.float32) / np.float32(255.0), msk[n].astype(np.float32) / np.float32(255.0), bool(f),
                            np.float32(a), np.float32(b), None)
        assert np.array_equal(oi[n, 0].cpu().numpy(), ri), n          # bit-exact: same two roundings as numpy
        assert np.array_equal(om[n, 0].cpu().numpy(), rm), n
    # noise: mid-gray image, sigma 0.1 -> (almost) no clipping; mean 0, std sigma; reproducible per seed, different across seeds/samples
    flat = torch.full((2, 256, 256), 128, dtype=torch.uint8).cuda()
    prm = torch.tensor([[0, 1.0, 0.0, 0.1], [0, 1.0, 0.0, 0.1]])
    a, _ = imageops.augment(flat, None, prm, seed=7)
    b, _ = imageops.augment(flat, None, prm, seed=7)
    c, _ = imageops.augment(flat, None, prm, seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a[0], a[1])
    d = (a - 128 / 255).double()
    assert abs(float(d.mean())) < 2e-3 and abs(float(d.std()) - 0.1) < 2e-3
    assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    # the sampled parameters follow the reference pipeline's probabilities
    p = imageops.sample_augment_params(20000, torch.Generator().manual_seed(0))
    assert abs(float(p[:, 0].mean()) - 0.5) < 0.02
    assert abs(float(((p[:, 1] != 1) | (p[:, 2] != 0)).float().mean()) - 0.4) < 0.02
    assert abs(float((p[:, 3] > 0).float().mean()) - 0.1) < 0.01
    assert float(p[:, 1].min()) >= 0.8 and float(p[:, 1].max()) <= 1.2 and float(p[:, 3].max()) <= 0.44


def test_infer_image_equals_the_service_path(best_sd):
    """api/app.py:136-203 with the network in the middle: PIL gray + LANCZOS -> /255 -> net -> clip*255 -> uint8 -> LANCZOS back."""
    from image_enhancement_deglaring_b200 import imageops
    Image = pytest.importorskip("PIL.Image")
    net = dg.LightweightUNet(storage="fp16")
    net.load_state_dict(best_sd, strict=True)
    net = net.cuda().eval()
    rgb = _img(390, 610, 3, 11)
    got = imageops.infer_image(net, torch.from_numpy(rgb).cuda())[0].cpu().numpy()
    assert got.shape == (390, 610)
    gray = np.array(Image.fromarray(rgb).convert("L").resize((512, 512), Image.LANCZOS))
    with torch.no_grad():
        y = net.forward_u8(torch.from_numpy(gray)[None, None].cuda())[0, 0].cpu().numpy()
    want = np.array(Image.fromarray(y, mode="L").resize((610, 390), Image.LANCZOS))
    assert np.array_equal(got, want)
    from image_enhancement_deglaring_b200.session import InferenceSession
    assert np.array_equal(InferenceSession(net).infer_image(rgb), want)                      # the /infer handler's call
    assert np.array_equal(InferenceSession(net).infer_image(gray), np.array(Image.fromarray(y, mode="L")))   # 512x512 L upload


def test_errors():
    from image_enhancement_deglaring_b200 import imageops
    with pytest.raises(RuntimeError):
        imageops.pil_resize(torch.zeros(8, 8, dtype=torch.uint8), 4, 4)                 # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        imageops.pil_resize(torch.zeros(1, 8, 8, 2, dtype=torch.uint8).cuda(), 4, 4)    # 2 channels
    with pytest.raises(RuntimeError):
        imageops.cv2_resize(torch.zeros(1, 8, 8, 4, dtype=torch.uint8).cuda(), 4, 4)    # RGBA: cv2 path takes 1 or 3
