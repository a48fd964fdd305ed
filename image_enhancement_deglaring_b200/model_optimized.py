"""Drop-in `OptimizedUNet` (src/optimized_model.py:6-158) on the B200 fused ops.

Same constructor (`in_channels=1, out_channels=1`), the same 76 state_dict keys / shapes (parameter containers only), and
`forward(x[N,in,H,W]) -> [N,out,H,W]`.  No checkpoint ships for this architecture; parity is taken on deterministic
weights against the reference module (tests/golden/opt_rand.npz forward, opt_train.npz one training step).  Training: under
autograd the forward keeps the raw activations and `_backward` mirrors it op by op (per-op backward entry points of
include/deglare.h).

Mapping onto the fused 3x3 conv (include/deglare.h):
  _block conv .0 / .3          -> DG_X_SAME / DG_X_POOL2 sources (GroupNorm+SiLU[+AvgPool] on load)       :91-98, :33-42
  _upblock (nearest x2 + conv) -> DG_X_UP2 source; its own GroupNorm(4)+SiLU is applied by the consumer   :111-116
  ChannelAttention             -> global mean from the `act_sum` epilogue of the conv that pools the same tensor,
                                  two tiny mat-vecs (dg_channel_attention), scale applied on load (dg_src.scale) :185-202
  torch.cat((up, skip*att))    -> two-source conv, never materialised                                    :140-156
"""
import contextlib

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import DTYPE_CODES


def _on_device_of(t):
    """Make t's GPU the current one for the launches below (the public forward has already refused CPU tensors)."""
    return torch.cuda.device(t.device) if t.is_cuda else contextlib.nullcontext()


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
        dt = DTYPE_CODES[self.storage]
        pk = {}

        pk["_convs"] = []

        def conv(name, m):
            w = ops.pack_conv3x3(m.weight)
            pk[name] = (w, ops.pack_conv3x3_tc(w, dt) if self.path != 1 else None)
            pk["_convs"].append(name)

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

    def _packs_bwd(self):
        """Backward-only weight copies: taps-flipped [3,3,Co,Ci] fp32 for the CUDA-core data gradient and, in the 16-bit tiers,
        the bf16 tensor-core packing of the forward weights the tensor-core data gradient reads transposed."""
        pk = self._packs()
        if pk.get("_bwd") is None:
            bk = {}
            for name in pk["_convs"]:          # every 3x3 conv: "enc1.0" ... "dec1.3", "upconv4.1" ... "upconv1.1"
                v = pk[name]
                blk, idx = name.rsplit(".", 1)
                w = getattr(self, blk)[int(idx)].weight
                tc = tcf = None
                wflip = ops.flip_conv3x3(w)
                if self.storage != "fp32" and (self.path & 3) != 1:
                    tc = ops.pack_conv3x3_tc(v[0], ops.DG_BF16)
                    tcf = ops.pack_conv3x3_tc(wflip, ops.DG_BF16)   # for the tcgen05 data gradient of the pairs mma.sync lacks
                bk[name] = (wflip, tc, tcf)
            pk["_bwd"] = bk
        return pk["_bwd"]

    def forward(self, x):
        _lib.load()
        if not x.is_cuda:
            raise RuntimeError("OptimizedUNet (B200) runs on CUDA tensors only; there is no CPU fallback")
        if self.output.weight.device.type != "cuda":
            raise RuntimeError("OptimizedUNet (B200) needs its parameters on a CUDA device; there is no CPU fallback")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input [N,{self.in_channels},H,W], got {tuple(x.shape)}")
        N, _, H, W = x.shape
        if H % 16 or W % 16 or H < 16 or W < 16:
            raise RuntimeError(f"input {N}x{H}x{W}: H and W must be positive multiples of 16")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if x.requires_grad:
                raise NotImplementedError("gradient w.r.t. the input image is not implemented (the reference never needs it)")
            if getattr(self, "_dg_direct_grads", False):
                # train.GraphedTrainStep: inside a captured step the parameters are not autograd inputs (their AccumulateGrad nodes
                # would join the stream they were created on into the capture); a fresh leaf anchors the node and the gradients go
                # straight into the optimizer's flat bucket
                return _OptimizedUNetFn.apply(self, x, torch.zeros((), device=x.device, requires_grad=True))
            return _OptimizedUNetFn.apply(self, x, *self.parameters())
        with _on_device_of(x):
            return self._run(x)

    def _run(self, x, keep=None):
        """The forward of src/optimized_model.py:118-158 on the fused ops.  `keep` (training): filled with everything backward
        re-reads -- per conv its raw output, statistics and source descriptions, per attention its squeeze sums and scales."""
        N, _, H, W = x.shape
        x = x.detach().float().contiguous()
        pk = self._packs()
        dt = DTYPE_CODES[self.storage]
        dev = x.device

        def src(t, gnname, xform=ops.DG_X_SAME, scale=None):
            raw, stats, c = t
            g, b, groups = pk[gnname]
            return ops.make_src(raw, c, xform=xform, stats=stats, gamma=g, beta=b, groups=groups, scale=scale)

        def conv(name, srcs, cout, h, w, act_sum=None):
            wp, wtc = pk[name]
            raw, stats = ops.conv3x3_fused(srcs, wp, cout, N, h, w, dt, act_sum=act_sum, path=self.path, weight_tc=wtc)
            if keep is not None:
                keep[name] = dict(raw=raw, stats=stats, c=cout, h=h, w=w, srcs=srcs, cin=int(wp.shape[2]))
            return raw, stats, cout

        def attention(name, t, act_sum, h, w):
            w1, w2 = pk[name]
            scale = ops.channel_attention(act_sum, float(h * w), w1, w2)
            if keep is not None:
                keep[name] = dict(act_sum=act_sum, scale=scale, plane=float(h * w))
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
        hsrc = src(d, dname)
        if keep is not None:
            keep["head"] = hsrc
        return ops.head1x1(hsrc, hw_, hb, N, H, W, dt)

    # ---- training: autograd of the forward above (optimized_train.py:210/226 drives it through loss.backward()) ----------------
    _tc_dgrad_ok = {}   # (cin, cout) -> does the mma.sync data gradient cover the pair (probed once per process)
    _t5_dgrad_ok = {}   # (cin, cout, h, w) -> does the tcgen05 data gradient have a plan

    def _backward(self, keep, grad_y, flat):
        """All 76 parameter gradients into `flat` (zero on entry; parameters() order, the parameters' own layouts)."""
        pk, bk = self._packs(), self._packs_bwd()
        dt = DTYPE_CODES[self.storage]
        N, _, H, W = grad_y.shape
        dev = grad_y.device
        views, off = {}, 0
        for name, p in self.named_parameters():
            views[name] = flat[off:off + p.numel()]
            off += p.numel()
        T = {}   # conv name -> gradient at the conv's input grid, fp32 NHWC [N,h,w,cin_total]

        def gn_of(name):
            blk, idx = name.rsplit(".", 1)
            return f"{blk}.{int(idx) + 1}"

        def dgrad(name, dR):
            L = keep[name]
            wflip, wtc, wtcf = bk[name]
            key = (L["cin"], L["c"])
            if wtc is not None and OptimizedUNet._tc_dgrad_ok.get(key, True):
                try:
                    return ops.conv3x3_dgrad(dR, wtc, L["cin"], L["c"])
                except RuntimeError as e:   # rc 3 is raised before any launch: the pair is outside the tensor-core kernel's table
                    if "no tensor-core kernel" not in str(e):
                        raise
                    OptimizedUNet._tc_dgrad_ok[key] = False
            wkey = key + (L["h"], L["w"])
            if wtcf is not None and OptimizedUNet._t5_dgrad_ok.get(wkey, True):
                try:
                    return ops.conv3x3_dgrad_wide(dR, wtcf, L["cin"], L["c"])
                except RuntimeError as e:   # likewise: no plan for this shape, nothing was launched
                    if "no tcgen05 plan" not in str(e):
                        raise
                    OptimizedUNet._t5_dgrad_ok[wkey] = False
            return ops.conv3x3_dgrad_generic(dR, wflip, L["cin"], N, L["h"], L["w"])

        def finish(name, G, P):
            """G = dL/dy of conv `name` -> dR in place; GroupNorm / conv parameter gradients; the conv's input gradient."""
            L = keep[name]
            gn = gn_of(name)
            g, _, groups = pk[gn]
            ops.gn_bwd_apply(L["raw"], L["stats"], g, groups, dt, N, L["h"], L["w"], L["c"], P, G, views[gn + ".weight"],
                             views[gn + ".bias"])
            ops.conv3x3_wgrad(L["srcs"], G, L["cin"], L["c"], N, L["h"], L["w"], dt, path=self.path & 3,
                              out=views[name + ".weight"])
            if name != "enc1.0":
                T[name] = dgrad(name, G)

        def act(name, dA, off_a=0):
            L = keep[name]
            g, b, groups = pk[gn_of(name)]
            G = torch.empty((N, L["h"], L["w"], L["c"]), dtype=torch.float32, device=dev)
            P = torch.zeros((N, L["c"], 2), dtype=torch.float64, device=dev)
            ops.act_bwd(L["raw"], L["stats"], g, b, groups, dt, N, L["h"], L["w"], L["c"], G, P, dA_a=dA, off_a=off_a)
            finish(name, G, P)

        # head (src/optimized_model.py:158) and the last conv's activation
        L = keep["dec1.3"]
        G = torch.empty((N, H, W, L["c"]), dtype=torch.float32, device=dev)
        P = torch.zeros((N, L["c"], 2), dtype=torch.float64, device=dev)
        ops.head1x1_bwd(keep["head"], pk["head"][0], grad_y, N, H, W, dt, G, P, views["output.weight"], views["output.bias"])
        finish("dec1.3", G, P)
        # decoder, top level first (:138-156 backwards)
        for k in (1, 2, 3, 4):
            act(f"dec{k}.0", T.pop(f"dec{k}.3"))
            act(f"upconv{k}.1", T[f"dec{k}.0"], 0)                    # first half of torch.cat((dec, enc * att))
            prod = f"dec{k + 1}.3" if k < 4 else "bottleneck.3"       # nn.Upsample(x2, nearest) backward: 2x2 sums
            Lp = keep[prod]
            act(prod, ops.grad_gather(N, Lp["h"], Lp["w"], Lp["c"], u=T.pop(f"upconv{k}.1")))
        act("bottleneck.0", T.pop("bottleneck.3"))
        # encoder: each skip tensor feeds the decoder concat (x att), the next level's pool and the attention squeeze
        for lvl in (4, 3, 2, 1):
            name = f"enc{lvl}.3"
            nxt = "bottleneck.0" if lvl == 4 else f"enc{lvl + 1}.0"
            L = keep[name]
            g, b, groups = pk[gn_of(name)]
            att = keep[f"attention{lvl}"]
            w1, w2 = pk[f"attention{lvl}"]
            tdec = T.pop(f"dec{lvl}.0")
            dscale = ops.scale_bwd_sum(L["raw"], L["stats"], g, b, groups, dt, N, L["h"], L["w"], L["c"], tdec, L["c"])
            add = ops.channel_attention_bwd(att["act_sum"], att["plane"], w1, w2, dscale, views[f"attention{lvl}.fc.0.weight"],
                                            views[f"attention{lvl}.fc.2.weight"])
            dA = ops.grad_gather(N, L["h"], L["w"], L["c"], a=tdec, off_a=L["c"], a_scale=att["scale"], b=T.pop(nxt), add=add)
            del tdec
            act(name, dA)
            del dA
            act(f"enc{lvl}.0", T.pop(name))


class _OptimizedUNetFn(torch.autograd.Function):
    """Autograd bridge of OptimizedUNet: forward keeps the raw activations, backward writes every parameter gradient into one
    flat buffer (the optimizer's own bucket when FusedAdamW / FlatGradBucket own `.grad` and it is fresh after zero_grad) and
    ends with the data-parallel mean all-reduce, exactly like the LightweightUNet bridge (train._LightweightUNetFn)."""

    @staticmethod
    def forward(ctx, module, x, *params):
        keep = {}
        with _on_device_of(x):
            y = module._run(x, keep)
        ctx.module, ctx.keep, ctx.n_inputs = module, keep, len(params)
        ctx.direct = getattr(module, "_dg_direct_grads", False)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        from .train import _flat_grad_sink, sync_gradients
        module, keep = ctx.module, ctx.keep
        if keep is None:
            raise RuntimeError("OptimizedUNet: backward through the same forward twice (the saved activations were released)")
        params = list(module.parameters())
        total = sum(p.numel() for p in params)
        grad_y = grad_y.detach().float().contiguous()
        with _on_device_of(grad_y):
            sink = _flat_grad_sink(params)
            if ctx.direct and sink is None:
                raise RuntimeError("GraphedTrainStep: the gradients must land in the optimizer's flat bucket (FusedAdamW.zero_grad "
                                   "inside the step, no per-parameter hooks)")
            flat = sink if sink is not None else torch.zeros(total, dtype=torch.float32, device=grad_y.device)
            module._backward(keep, grad_y, flat)
        ctx.keep = None
        sync_gradients(flat, getattr(module, "ddp_sync", True))
        if sink is not None:
            sink._dg_zero_version = None      # the bucket now holds a gradient: a second backward accumulates the ordinary way
            return (None, None) + (None,) * ctx.n_inputs
        grads, off = [], 0
        for p in params:
            n = p.numel()
            grads.append(flat[off:off + n].view_as(p) if p.requires_grad else None)
            off += n
        return (None, None, *grads)
