"""Torch (CPU) stand-ins of the per-op C-ABI calls OptimizedUNet's host orchestration makes -- TEST INFRASTRUCTURE.

`model_optimized.OptimizedUNet._run` / `._backward` are Python above the C-ABI: which kernel reads which tensor, with which
channel window, in which order.  That wiring is host logic and is tested without a GPU by swapping `ops.*` for the functions
below, each a direct restatement of the semantics `include/deglare.h` documents for the entry point of the same name
(never of the kernels' code).  The product never imports this file; the real kernels are tested on the GPU against the same
oracle (tests/test_gpu_optimized.py).
"""
import torch
import torch.nn.functional as F

from image_enhancement_deglaring_b200 import _lib

DG_F32, DG_F16, DG_BF16 = _lib.DG_F32, _lib.DG_F16, _lib.DG_BF16
TORCH_DTYPE = {DG_F32: torch.float32, DG_F16: torch.float16, DG_BF16: torch.bfloat16}
EPS = 1e-5


class Src:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_src(raw, channels, xform=_lib.DG_X_SAME, stats=None, gamma=None, beta=None, groups=1, silu=True, scale=None, **_):
    return Src(raw=raw, channels=channels, xform=xform, stats=stats, gamma=gamma, beta=beta, groups=groups, silu=silu, scale=scale)


def _mean_rstd(stats, groups, plane):
    """[N,C,2] (sum, sumsq) -> per-channel mean / rstd of the channel's group, [N,C] each."""
    N, C, _ = stats.shape
    g = stats.reshape(N, groups, C // groups, 2).sum(dim=2) / (plane * (C // groups))
    mean = g[..., 0]
    rstd = 1.0 / torch.sqrt((g[..., 1] - mean * mean).clamp_min(0) + EPS)
    rep = C // groups
    return mean.repeat_interleave(rep, 1).float(), rstd.repeat_interleave(rep, 1).float()


def _xhat_y(raw, stats, gamma, beta, groups):
    """raw NHWC -> (xhat, y) NHWC fp32."""
    r = raw.float()
    N, H, W, C = r.shape
    mean, rstd = _mean_rstd(stats, groups, H * W)
    xh = (r - mean[:, None, None, :]) * rstd[:, None, None, :]
    return xh, xh * gamma + beta, rstd


def _activated(s):
    """dg_src semantics: affine, SiLU, scale, then the spatial transform; NCHW fp32 on the consumer's grid."""
    if s.xform == _lib.DG_X_IMAGE:
        return s.raw.float()
    y = s.raw.float()
    if s.stats is not None:
        _, y, _ = _xhat_y(s.raw, s.stats, s.gamma, s.beta, s.groups)
    if s.silu:
        y = F.silu(y)
    if s.scale is not None:
        y = y * s.scale[:, None, None, :]
    y = y.permute(0, 3, 1, 2)
    if s.xform == _lib.DG_X_POOL2:
        y = F.avg_pool2d(y, 2, 2)
    elif s.xform == _lib.DG_X_UP2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    elif s.xform != _lib.DG_X_SAME:
        raise NotImplementedError(s.xform)
    return y


def _stats_of(raw):
    r = raw.double()
    return torch.stack((r.sum(dim=(1, 2)), (r * r).sum(dim=(1, 2))), dim=-1).contiguous()


def pack_conv3x3(w):
    return w.detach().float().permute(2, 3, 1, 0).contiguous()


def flip_conv3x3(w):
    return w.detach().float().flip(2, 3).permute(2, 3, 0, 1).contiguous()


def pack_conv3x3_tc(w_packed, dtype, stream=None):
    return None


def conv3x3_fused(srcs, weight, cout, N, H, W, dtype, out=None, out_stats=None, act_sum=None, path=0, stream=None, eps=EPS,
                  weight_tc=None, weight_comp=None, want_stats=True):
    A = torch.cat([_activated(s) for s in srcs], dim=1)
    assert A.shape == (N, weight.shape[2], H, W), (A.shape, weight.shape, H, W)
    raw = F.conv2d(A, weight.permute(3, 2, 0, 1), None, 1, 1).permute(0, 2, 3, 1).contiguous()
    assert raw.shape[-1] == cout
    if act_sum is not None:   # pixel sums of the ACTIVATED src[0] at its own resolution
        s0 = srcs[0]
        _, y, _ = _xhat_y(s0.raw, s0.stats, s0.gamma, s0.beta, s0.groups)
        act_sum += F.silu(y).double().sum(dim=(1, 2))
    stats = _stats_of(raw) if want_stats else None
    raw = raw.to(TORCH_DTYPE[dtype])
    if out is not None:
        out.copy_(raw)
        raw = out
    return raw, stats


def conv3x3_wgrad(srcs, dR, cin_total, cout, N, H, W, dtype, path=0, stream=None, eps=EPS, out=None):
    A = torch.cat([_activated(s) for s in srcs], dim=1)
    dW = torch.nn.grad.conv2d_weight(A, (cout, cin_total, 3, 3), dR.permute(0, 3, 1, 2).contiguous(), padding=1)
    if out is None:
        return dW
    out += dW.reshape(out.shape)
    return out


def conv3x3_dgrad_generic(dR, weight_flip, cin, N, H, W, out=None, stream=None):
    src = make_src(dR, int(dR.shape[-1]), silu=False)
    return conv3x3_fused([src], weight_flip, cin, N, H, W, DG_F32, out=out, path=1, want_stats=False)[0]


def conv3x3_dgrad(*a, **k):
    raise RuntimeError("libdeglare error 3: dgrad: no tensor-core kernel (stand-in)")


def head1x1(src, weight, bias, N, H, W, dtype, out=None, target=None, l1_sum=None, stream=None, eps=EPS):
    A = _activated(src)
    return F.conv2d(A, weight[:, :, None, None], bias)


def head1x1_bwd(src, weight, grad_y, N, H, W, dtype, G, P, dW, dB, stream=None, eps=EPS):
    xh, y, _ = _xhat_y(src.raw, src.stats, src.gamma, src.beta, src.groups)
    A = F.silu(y)
    dO = grad_y.permute(0, 2, 3, 1)                                  # [N,H,W,OC]
    dA = dO @ weight                                                 # [N,H,W,C]
    s = torch.sigmoid(y)
    g = dA * (s * (1 + y * (1 - s)))
    G.copy_(g)
    P[..., 0] += g.double().sum(dim=(1, 2))
    P[..., 1] += (g * xh).double().sum(dim=(1, 2))
    dW += torch.einsum("nhwo,nhwc->oc", dO, A).reshape(dW.shape)
    dB += dO.sum(dim=(0, 1, 2))


def act_bwd(raw, stats, gamma, beta, groups, dtype, N, H, W, channels, G, P, dA_a=None, off_a=0, dA_b=None, off_b=0, stream=None,
            eps=EPS):
    xh, y, _ = _xhat_y(raw, stats, gamma, beta, groups)
    d = torch.zeros_like(y)
    if dA_a is not None:
        assert dA_a.shape[:3] == (N, H, W)
        d = d + dA_a[..., off_a:off_a + channels]
    if dA_b is not None:
        assert dA_b.shape[:3] == (N, H // 2, W // 2)
        d = d + 0.25 * dA_b[..., off_b:off_b + channels].repeat_interleave(2, 1).repeat_interleave(2, 2)
    s = torch.sigmoid(y)
    g = d * (s * (1 + y * (1 - s)))
    G.copy_(g)
    P[..., 0] += g.double().sum(dim=(1, 2))
    P[..., 1] += (g * xh).double().sum(dim=(1, 2))


def gn_bwd_apply(raw, stats, gamma, groups, dtype, N, H, W, channels, P, G, dgamma, dbeta, stream=None, eps=EPS):
    xh, _, rstd = _xhat_y(raw, stats, gamma, torch.zeros_like(gamma), groups)
    cpg = channels // groups
    gp = (gamma.double()[None, :, None] * P).reshape(N, groups, cpg, 2).sum(dim=2) / (H * W * cpg)    # [N,groups,2]
    m = gp.repeat_interleave(cpg, 1).float()                                                          # [N,C,2]
    dR = rstd[:, None, None, :] * (gamma * G - m[:, None, None, :, 0] - xh * m[:, None, None, :, 1])
    G.copy_(dR)
    if dgamma is not None:
        dgamma += P[..., 1].sum(dim=0).float()
        dbeta += P[..., 0].sum(dim=0).float()


def grad_gather(N, H, W, channels, a=None, off_a=0, a_scale=None, b=None, off_b=0, u=None, off_u=0, add=None, out=None, stream=None):
    acc = torch.zeros(N, H, W, channels)
    if a is not None:
        assert a.shape[:3] == (N, H, W)
        t = a[..., off_a:off_a + channels]
        acc = acc + (t * a_scale[:, None, None, :] if a_scale is not None else t)
    if b is not None:
        assert b.shape[:3] == (N, H // 2, W // 2)
        acc = acc + 0.25 * b[..., off_b:off_b + channels].repeat_interleave(2, 1).repeat_interleave(2, 2)
    if u is not None:
        assert u.shape[:3] == (N, 2 * H, 2 * W)
        acc = acc + u[..., off_u:off_u + channels].reshape(N, H, 2, W, 2, channels).sum(dim=(2, 4))
    if add is not None:
        acc = acc + add[:, None, None, :]
    return acc.contiguous()


def scale_bwd_sum(raw, stats, gamma, beta, groups, dtype, N, H, W, channels, d, off_d, stream=None, eps=EPS):
    _, y, _ = _xhat_y(raw, stats, gamma, beta, groups)
    return (d[..., off_d:off_d + channels] * F.silu(y)).double().sum(dim=(1, 2))


def channel_attention(act_sum, plane, w1, w2, stream=None):
    m = (act_sum / plane).float()
    return torch.sigmoid(F.silu(m @ w1.t()) @ w2.t())


def channel_attention_bwd(act_sum, plane, w1, w2, dscale, dw1, dw2, stream=None):
    m = (act_sum / plane).float()
    a = m @ w1.t()
    h = F.silu(a)
    s = torch.sigmoid(h @ w2.t())
    dz = dscale.float() * s * (1 - s)
    dw2 += (dz.t() @ h).reshape(dw2.shape)
    dh = dz @ w2
    sa = torch.sigmoid(a)
    da = dh * (sa * (1 + a * (1 - sa)))
    dw1 += (da.t() @ m).reshape(dw1.shape)
    return (da @ w1) / plane


STANDINS = ("make_src", "pack_conv3x3", "flip_conv3x3", "pack_conv3x3_tc", "conv3x3_fused", "conv3x3_wgrad", "conv3x3_dgrad_generic",
            "conv3x3_dgrad", "head1x1", "head1x1_bwd", "act_bwd", "gn_bwd_apply", "grad_gather", "scale_bwd_sum",
            "channel_attention", "channel_attention_bwd")


def install(monkeypatch, ops_module):
    g = globals()
    for name in STANDINS:
        monkeypatch.setattr(ops_module, name, g[name])
