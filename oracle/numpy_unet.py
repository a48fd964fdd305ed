"""CPU oracle (numpy) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy restatement of the arithmetic on the reference's UNet de-glaring hot
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package; the product
(`image_enhancement_deglaring_b200`) never does and fails loudly without its CUDA
library.

Parity status: PINNED.  The reference has no golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference module
itself, executed in the build container by `tests/golden/make_golden.py` (which
imports /root/reference/src/model.py and src/optimized_model.py by path) and
committed under `tests/golden/`.  `tests/test_oracle.py` re-checks the oracle against
those fixtures on every run.

All arithmetic of the reference lives in torch==2.7.0 ATen operators (un-vendored);
the operator semantics restated here are the published PyTorch ones:
  Conv2d(k=3, padding=1, bias=False)            src/model.py:93,96
  GroupNorm(G, C), eps=1e-5, biased variance     src/model.py:94,97
  SiLU  x*sigmoid(x)                             src/model.py:95,98
  AvgPool2d(2, 2)                                src/model.py:35-41
  ConvTranspose2d(2C, C, k=2, s=2) weight [Cin,Cout,2,2] + bias   src/model.py:47-53
  torch.cat((up, skip), dim=1)                   src/model.py:116,120,124,128
  Conv2d(C, out, k=1) + bias                     src/model.py:57,131
  Upsample(scale_factor=2, mode='nearest')       src/optimized_model.py:112
  ChannelAttention (SE)                          src/optimized_model.py:161-202

Layout here is NCHW like the reference; dtype is selectable (float32 mirrors the
reference, float64 gives a rounding-free yardstick).
"""
import numpy as np

EPS = 1e-5


def conv3x3(x, w):
    """3x3 stride-1 zero-pad-1 cross-correlation, no bias.  x [N,Ci,H,W], w [Co,Ci,3,3]."""
    n, ci, h, wd = x.shape
    co = w.shape[0]
    xp = np.zeros((n, ci, h + 2, wd + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((n, co, h, wd), dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            win = xp[:, :, ky:ky + h, kx:kx + wd]
            out += np.einsum("nchw,oc->nohw", win, w[:, :, ky, kx], optimize=True)
    return out


def conv1x1(x, w, b):
    """1x1 conv + bias.  w [Co,Ci,1,1], b [Co]."""
    return np.einsum("nchw,oc->nohw", x, w[:, :, 0, 0], optimize=True) + b[None, :, None, None]


def group_norm(x, groups, gamma, beta, eps=EPS):
    """torch.nn.GroupNorm: per (n, group) mean / biased variance over (C/G, H, W)."""
    n, c, h, w = x.shape
    xg = x.reshape(n, groups, -1)
    mu = xg.mean(axis=2, keepdims=True)
    var = ((xg - mu) ** 2).mean(axis=2, keepdims=True)
    y = ((xg - mu) / np.sqrt(var + eps)).reshape(n, c, h, w)
    return y * gamma[None, :, None, None] + beta[None, :, None, None]


def silu(x):
    return x / (1.0 + np.exp(-x))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def avg_pool2(x):
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).mean(axis=(3, 5))


def conv_transpose2x2(x, w, b):
    """ConvTranspose2d(k=2, s=2): up[n,co,2i+a,2j+b] = bias[co] + sum_ci x[n,ci,i,j] w[ci,co,a,b]."""
    n, ci, h, wd = x.shape
    co = w.shape[1]
    out = np.empty((n, co, 2 * h, 2 * wd), dtype=x.dtype)
    for a in range(2):
        for bb in range(2):
            out[:, :, a::2, bb::2] = np.einsum("nchw,co->nohw", x, w[:, :, a, bb], optimize=True)
    return out + b[None, :, None, None]


def upsample_nearest2(x):
    return x.repeat(2, axis=2).repeat(2, axis=3)


def _groups_lightweight(features, num_groups):
    """Group-count rule of LightweightUNet._block (src/model.py:69-86)."""
    g = num_groups
    if features < num_groups:
        g = features
        for i in range(min(num_groups, features), 0, -1):
            if features % i == 0:
                g = i
                break
    elif features % num_groups != 0:
        for i in range(num_groups, 0, -1):
            if features % i == 0:
                g = i
                break
    return g


def _block(x, sd, prefix, groups, taps=None):
    """Conv-GN-SiLU-Conv-GN-SiLU (src/model.py:92-99).  `taps` collects raw conv outputs."""
    for conv, gn in (("0", "1"), ("3", "4")):
        raw = conv3x3(x, sd[f"{prefix}.{conv}.weight"])
        if taps is not None:
            taps[f"{prefix}.{conv}"] = raw
        x = silu(group_norm(raw, groups, sd[f"{prefix}.{gn}.weight"], sd[f"{prefix}.{gn}.bias"]))
    return x


def lightweight_unet_forward(x, sd, num_groups=8, taps=None):
    """LightweightUNet.forward (src/model.py:101-133).  sd: {key: ndarray}, x [N,Cin,H,W]."""
    dt = x.dtype
    sd = {k: np.asarray(v, dtype=dt) for k, v in sd.items()}
    g = lambda pfx: _groups_lightweight(sd[f"{pfx}.0.weight"].shape[0], num_groups)
    enc1 = _block(x, sd, "enc1", g("enc1"), taps)
    enc2 = _block(avg_pool2(enc1), sd, "enc2", g("enc2"), taps)
    enc3 = _block(avg_pool2(enc2), sd, "enc3", g("enc3"), taps)
    enc4 = _block(avg_pool2(enc3), sd, "enc4", g("enc4"), taps)
    d = _block(avg_pool2(enc4), sd, "bottleneck", g("bottleneck"), taps)
    for lvl, skip in ((4, enc4), (3, enc3), (2, enc2), (1, enc1)):
        up = conv_transpose2x2(d, sd[f"upconv{lvl}.weight"], sd[f"upconv{lvl}.bias"])
        if taps is not None:
            taps[f"upconv{lvl}"] = up
        d = _block(np.concatenate((up, skip), axis=1), sd, f"dec{lvl}", g(f"dec{lvl}"), taps)
    return conv1x1(d, sd["output_conv.weight"], sd["output_conv.bias"])


def _channel_attention(x, sd, prefix):
    """ChannelAttention.forward (src/optimized_model.py:185-202)."""
    avg = x.mean(axis=(2, 3))
    h = silu(avg @ sd[f"{prefix}.fc.0.weight"].T)
    wgt = sigmoid(h @ sd[f"{prefix}.fc.2.weight"].T)
    return x * wgt[:, :, None, None]


def _upblock(x, sd, prefix):
    """_upblock: nearest x2 -> Conv3x3 -> GN(4) -> SiLU (src/optimized_model.py:111-116)."""
    raw = conv3x3(upsample_nearest2(x), sd[f"{prefix}.1.weight"])
    return silu(group_norm(raw, 4, sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"]))


def optimized_unet_forward(x, sd):
    """OptimizedUNet.forward (src/optimized_model.py:118-158)."""
    dt = x.dtype
    sd = {k: np.asarray(v, dtype=dt) for k, v in sd.items()}
    grp = lambda pfx, g: max(1, min(g, sd[f"{pfx}.0.weight"].shape[0]))
    enc1 = _block(x, sd, "enc1", grp("enc1", 1))
    enc2 = _block(avg_pool2(enc1), sd, "enc2", grp("enc2", 4))
    enc3 = _block(avg_pool2(enc2), sd, "enc3", grp("enc3", 4))
    enc4 = _block(avg_pool2(enc3), sd, "enc4", grp("enc4", 4))
    d = _block(avg_pool2(enc4), sd, "bottleneck", 8)
    for lvl, skip in ((4, enc4), (3, enc3), (2, enc2), (1, enc1)):
        up = _upblock(d, sd, f"upconv{lvl}")
        skip = _channel_attention(skip, sd, f"attention{lvl}")
        d = _block(np.concatenate((up, skip), axis=1), sd, f"dec{lvl}", grp(f"dec{lvl}", 4))
    return conv1x1(d, sd["output.weight"], sd["output.bias"])


def l1_loss(out, target):
    """nn.L1Loss() mean reduction (optimized_train.py:439)."""
    return np.abs(out - target).mean()


def l1_loss_grad(out, target):
    """d mean|o-t| / d o = sign(o-t)/numel, sign(0) = 0."""
    return np.sign(out - target) / out.size
