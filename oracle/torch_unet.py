"""CPU oracle (torch functional, fp32/fp64) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The same restatement as `oracle/numpy_unet.py`, written with torch.nn.functional so
that (a) it is fast enough to time as the CPU baseline on all host cores and (b)
autograd gives the training-step oracle (gradients of every parameter, clip, AdamW).
It deliberately does NOT import the reference (`/root/reference` is absent on the GPU
box) and is never imported by the product package.

Parity status: PINNED against the reference modules executed in the build container
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`, checked by
`tests/test_oracle.py`) and cross-checked against the independent numpy restatement.

Reference anchors:
  LightweightUNet.forward          src/model.py:101-133      -> lightweight_forward
  LightweightUNet._block           src/model.py:59-99        -> _block
  OptimizedUNet.forward            src/optimized_model.py:118-158 -> optimized_forward
  ChannelAttention.forward         src/optimized_model.py:185-202 -> _channel_attention
  plain fp32 training branch       optimized_train.py:220-233 -> train_step
  L1Loss / AdamW / clip            optimized_train.py:439-446, :230
"""
import math

import torch
import torch.nn.functional as F

EPS = 1e-5


def groups_lightweight(features, num_groups):
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


def gn_silu(raw, groups, gamma, beta):
    return F.silu(F.group_norm(raw, groups, gamma, beta, EPS))


def _block(x, sd, prefix, groups, taps=None):
    for conv, gn in (("0", "1"), ("3", "4")):
        raw = F.conv2d(x, sd[f"{prefix}.{conv}.weight"], None, 1, 1)
        if taps is not None:
            taps[f"{prefix}.{conv}"] = raw
        x = gn_silu(raw, groups, sd[f"{prefix}.{gn}.weight"], sd[f"{prefix}.{gn}.bias"])
    return x


def lightweight_forward(x, sd, num_groups=8, taps=None):
    """x [N,Cin,H,W]; sd: state_dict of tensors; returns [N,Cout,H,W] (no clip, no sigmoid)."""
    g = lambda pfx: groups_lightweight(sd[f"{pfx}.0.weight"].shape[0], num_groups)
    enc1 = _block(x, sd, "enc1", g("enc1"), taps)
    enc2 = _block(F.avg_pool2d(enc1, 2, 2), sd, "enc2", g("enc2"), taps)
    enc3 = _block(F.avg_pool2d(enc2, 2, 2), sd, "enc3", g("enc3"), taps)
    enc4 = _block(F.avg_pool2d(enc3, 2, 2), sd, "enc4", g("enc4"), taps)
    d = _block(F.avg_pool2d(enc4, 2, 2), sd, "bottleneck", g("bottleneck"), taps)
    for lvl, skip in ((4, enc4), (3, enc3), (2, enc2), (1, enc1)):
        up = F.conv_transpose2d(d, sd[f"upconv{lvl}.weight"], sd[f"upconv{lvl}.bias"], stride=2)
        if taps is not None:
            taps[f"upconv{lvl}"] = up
        d = _block(torch.cat((up, skip), dim=1), sd, f"dec{lvl}", g(f"dec{lvl}"), taps)
    return F.conv2d(d, sd["output_conv.weight"], sd["output_conv.bias"])


def _channel_attention(x, sd, prefix):
    avg = x.mean(dim=(2, 3))
    h = F.silu(avg @ sd[f"{prefix}.fc.0.weight"].t())
    w = torch.sigmoid(h @ sd[f"{prefix}.fc.2.weight"].t())
    return x * w[:, :, None, None]


def _upblock(x, sd, prefix, taps=None):
    raw = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), sd[f"{prefix}.1.weight"], None, 1, 1)
    if taps is not None:
        taps[f"{prefix}.1"] = raw
    return gn_silu(raw, 4, sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"])


def optimized_forward(x, sd, taps=None):
    grp = lambda pfx, g: max(1, min(g, sd[f"{pfx}.0.weight"].shape[0]))
    enc1 = _block(x, sd, "enc1", grp("enc1", 1), taps)
    enc2 = _block(F.avg_pool2d(enc1, 2, 2), sd, "enc2", grp("enc2", 4), taps)
    enc3 = _block(F.avg_pool2d(enc2, 2, 2), sd, "enc3", grp("enc3", 4), taps)
    enc4 = _block(F.avg_pool2d(enc3, 2, 2), sd, "enc4", grp("enc4", 4), taps)
    d = _block(F.avg_pool2d(enc4, 2, 2), sd, "bottleneck", 8, taps)
    for lvl, skip in ((4, enc4), (3, enc3), (2, enc2), (1, enc1)):
        up = _upblock(d, sd, f"upconv{lvl}", taps)
        skip = _channel_attention(skip, sd, f"attention{lvl}")
        d = _block(torch.cat((up, skip), dim=1), sd, f"dec{lvl}", grp(f"dec{lvl}", 4), taps)
    return F.conv2d(d, sd["output.weight"], sd["output.bias"])


def l1_loss(out, target):
    return (out - target).abs().mean()


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (L2): coef = min(1, max_norm / (total + 1e-6))."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in grads.items()}, float(total)


def adamw_step(params, grads, state, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.AdamW single-tensor math (decoupled weight decay), in place on copies."""
    b1, b2 = betas
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    out = {}
    for k, p in params.items():
        g = grads[k]
        m = state.setdefault(("m", k), torch.zeros_like(p))
        v = state.setdefault(("v", k), torch.zeros_like(p))
        p = p * (1.0 - lr * weight_decay)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        out[k] = p - (lr / bc1) * (m / denom)
    return out


def train_step(sd, x, target, forward=lightweight_forward, lr=0.002362532125818593,
               weight_decay=6.753784966611083e-05, max_norm=1.0, state=None):
    """One optimisation step of optimized_train.py:220-233 (fp32 branch).

    Returns dict(loss, grads (pre-clip), total_norm, new_params)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = forward(x, params)
    loss = l1_loss(out, target)
    gl = torch.autograd.grad(loss, list(params.values()))
    grads = dict(zip(params.keys(), gl))
    clipped, total = clip_grad_norm(grads, max_norm) if max_norm > 0 else (grads, 0.0)
    state = {} if state is None else state
    new = adamw_step({k: v.detach() for k, v in params.items()}, clipped, state, lr, weight_decay)
    return {"loss": float(loss.detach()), "out": out.detach(), "grads": grads, "total_norm": total,
            "new_params": new, "state": state}
