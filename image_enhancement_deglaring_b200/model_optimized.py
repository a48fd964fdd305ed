"""Drop-in `OptimizedUNet` (src/optimized_model.py:6-158) on the B200 fused ops.

Same constructor (`in_channels=1, out_channels=1`), the same 76 state_dict keys / shapes (parameter containers only), and
`forward(x[N,in,H,W]) -> [N,out,H,W]`.  No checkpoint ships for this architecture; parity is taken on deterministic
weights against the reference module (tests/golden/opt_rand.npz).  Inference only for now.

Mapping onto the fused 3x3 conv (include/deglare.h):
  _block conv .0 / .3          -> DG_X_SAME / DG_X_POOL2 sources (GroupNorm+SiLU[+AvgPool] on load)       :91-98, :33-42
  _upblock (nearest x2 + conv) -> DG_X_UP2 source; its own GroupNorm(4)+SiLU is applied by the consumer   :111-116
  ChannelAttention             -> global mean from the `act_sum` epilogue of the conv that pools the same tensor,
                                  two tiny mat-vecs (dg_channel_attention), scale applied on load (dg_src.scale) :185-202
  torch.cat((up, skip*att))    -> two-source conv, never materialised                                    :140-156
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import DTYPE_CODES


class ChannelAttention(nn.Module):
    """Parameter container with the reference's layout (src/optimized_model.py:161-183); evaluated by dg_channel_attention."""

    def __init__(self, channels, reduction=16):
        super().__init__()
        reduced = max(channels // reduction, 8)
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channels, reduced, bias=False), nn.SiLU(inplace=True),
                                nn.Linear(reduced, channels, bias=False), nn.Sigmoid())


class OptimizedUNet(nn.Module):
    def __init__(self, in_channels=1, out_channels=1, *, storage="fp32", path=0):
        super().__init__()
        if storage not in DTYPE_CODES:
            raise ValueError(f"storage must be one of {sorted(DTYPE_CODES)}")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.storage, self.path = storage, path
        f = self.init_features = 16
        self.enc1 = self._block(in_channels, f, groups=1)
        self.pool1 = nn.AvgPool2d(kernel_size=2, stride=2)
        self.enc2 = self._block(f, f * 2)
        self.pool2 = nn.AvgPool2d(kernel_size=2, stride=2)
        self.enc3 = self._block(f * 2, f * 4)
        self.pool3 = nn.AvgPool2d(kernel_size=2, stride=2)
        self.enc4 = self._block(f * 4, f * 8)
        self.pool4 = nn.AvgPool2d(kernel_size=2, stride=2)
        self.bottleneck = nn.Sequential(
            nn.Conv2d(f * 8, f * 16, kernel_size=3, padding=1, bias=False), nn.GroupNorm(8, f * 16), nn.SiLU(inplace=True),
            nn.Conv2d(f * 16, f * 16, kernel_size=3, padding=1, bias=False), nn.GroupNorm(8, f * 16), nn.SiLU(inplace=True))
        self.attention4 = ChannelAttention(f * 8)
        self.attention3 = ChannelAttention(f * 4)
        self.attention2 = ChannelAttention(f * 2)
        self.attention1 = ChannelAttention(f)
        self.upconv4 = self._upblock(f * 16, f * 8)
        self.dec4 = self._block(f * 16, f * 8)
        self.upconv3 = self._upblock(f * 8, f * 4)
        self.dec3 = self._block(f * 8, f * 4)
        self.upconv2 = self._upblock(f * 4, f * 2)
        self.dec2 = self._block(f * 4, f * 2)
        self.upconv1 = self._upblock(f * 2, f)
        self.dec1 = self._block(f * 2, f)
        self.output = nn.Conv2d(f, out_channels, kernel_size=1)
        self._pack_key = None
        self._pk = None

    @staticmethod
    def _block(in_channels, features, groups=4):
        groups = max(1, min(groups, features))
        return nn.Sequential(
            nn.Conv2d(in_channels, features, kernel_size=3, padding=1, bias=False), nn.GroupNorm(groups, features),
            nn.SiLU(inplace=True),
            nn.Conv2d(features, features, kernel_size=3, padding=1, bias=False), nn.GroupNorm(groups, features),
            nn.SiLU(inplace=True))

    @staticmethod
    def _upblock(in_channels, out_channels):
        return nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"),
                             nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, bias=False),
                             nn.GroupNorm(4, out_channels), nn.SiLU(inplace=True))

    # ---- packed parameters -------------------------------------------------------------------------------------------------
    def _packs(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (self.storage, self.path, _lib.generation())
        if key == self._pack_key:
            return self._pk
        if self.output.weight.device.type != "cuda":
            raise RuntimeError("OptimizedUNet (B200) needs its parameters on a CUDA device; there is no CPU fallback")
        dt = DTYPE_CODES[self.storage]
        pk = {}

        def conv(name, m):
            w = ops.pack_conv3x3(m.weight)
            pk[name] = (w, ops.pack_conv3x3_tc(w, dt) if self.path != 1 else None)

        def gn(name, m):
            pk[name] = (m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous(), m.num_groups)

        for b in ("enc1", "enc2", "enc3", "enc4", "bottleneck", "dec4", "dec3", "dec2", "dec1"):
            blk = getattr(self, b)
            conv(b + ".0", blk[0]); gn(b + ".1", blk[1]); conv(b + ".3", blk[3]); gn(b + ".4", blk[4])
        for u in ("upconv4", "upconv3", "upconv2", "upconv1"):
            blk = getattr(self, u)
            conv(u + ".1", blk[1]); gn(u + ".2", blk[2])
        for a in ("attention4", "attention3", "attention2", "attention1"):
            fc = getattr(self, a).fc
            pk[a] = (fc[0].weight.detach().float().contiguous(), fc[2].weight.detach().float().contiguous())
        pk["head"] = (self.output.weight.detach().float().reshape(self.out_channels, -1).contiguous(),
                      self.output.bias.detach().float().contiguous())
        self._pk, self._pack_key = pk, key
        return pk

    def forward(self, x):
        lib = _lib.load()
        if not x.is_cuda:
            raise RuntimeError("OptimizedUNet (B200) runs on CUDA tensors only; there is no CPU fallback")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input [N,{self.in_channels},H,W], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("OptimizedUNet training is not implemented yet: call under torch.no_grad()")
        N, _, H, W = x.shape
        if H % 16 or W % 16 or H < 16 or W < 16:
            raise RuntimeError(f"input {N}x{H}x{W}: H and W must be positive multiples of 16")
        x = x.detach().float().contiguous()
        pk = self._packs()
        dt = DTYPE_CODES[self.storage]
        dev = x.device
        stream = torch.cuda.current_stream().cuda_stream

        def src(t, gnname, xform=ops.DG_X_SAME, scale=None):
            raw, stats, c = t
            g, b, groups = pk[gnname]
            return ops.make_src(raw, c, xform=xform, stats=stats, gamma=g, beta=b, groups=groups, scale=scale)

        def conv(name, srcs, cout, h, w, act_sum=None):
            wp, wtc = pk[name]
            raw, stats = ops.conv3x3_fused(srcs, wp, cout, N, h, w, dt, act_sum=act_sum, path=self.path, weight_tc=wtc)
            return raw, stats, cout

        def attention(name, t, act_sum, h, w):
            w1, w2 = pk[name]
            c = t[2]
            scale = torch.empty((N, c), dtype=torch.float32, device=dev)
            _lib.check(lib.dg_channel_attention(act_sum.data_ptr(), float(h * w), w1.data_ptr(), w2.data_ptr(), N, c,
                                                w1.shape[0], scale.data_ptr(), stream))
            return scale

        f = self.init_features
        chans = [f, f * 2, f * 4, f * 8, f * 16]
        hw = [(H >> i, W >> i) for i in range(5)]
        names = ["enc1", "enc2", "enc3", "enc4", "bottleneck"]
        # encoder: the pooling conv of the next block also reduces the activated skip's global mean (SE squeeze)
        t = conv("enc1.0", [ops.make_src(x, self.in_channels, xform=ops.DG_X_IMAGE, silu=False)], chans[0], *hw[0])
        t = conv("enc1.3", [src(t, "enc1.1")], chans[0], *hw[0])
        skips, scales = [t], []
        for lvl in range(1, 5):
            prev = names[lvl - 1]
            asum = torch.zeros((N, chans[lvl - 1]), dtype=torch.float64, device=dev)
            t = conv(names[lvl] + ".0", [src(skips[-1], prev + ".4", xform=ops.DG_X_POOL2)], chans[lvl], *hw[lvl], act_sum=asum)
            scales.append(attention(f"attention{lvl}", skips[-1], asum, *hw[lvl - 1]))
            t = conv(names[lvl] + ".3", [src(t, names[lvl] + ".1")], chans[lvl], *hw[lvl])
            if lvl < 4:
                skips.append(t)
        # decoder
        d, dname = t, "bottleneck.4"
        for lvl in (3, 2, 1, 0):
            k = lvl + 1
            up = conv(f"upconv{k}.1", [src(d, dname, xform=ops.DG_X_UP2)], chans[lvl], *hw[lvl])
            cat = [src(up, f"upconv{k}.2"), src(skips[lvl], names[lvl] + ".4", scale=scales[lvl])]
            d = conv(f"dec{k}.0", cat, chans[lvl], *hw[lvl])
            d = conv(f"dec{k}.3", [src(d, f"dec{k}.1")], chans[lvl], *hw[lvl])
            dname = f"dec{k}.4"
        hw_, hb = pk["head"]
        return ops.head1x1(src(d, dname), hw_, hb, N, H, W, dt)
