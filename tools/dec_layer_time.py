#!/usr/bin/env python
"""Times upconv1 + dec1.0 (16 -> 8 at 512x512, batch 64, fp16) through dg_conv3x3_fused: composite kernel vs round-1 fused kernel."""
import sys, os, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image_enhancement_deglaring_b200 import ops
N, H, W, c = 64, 512, 512, 8
dt = ops.DG_F16
rs = np.random.RandomState(0)
low = torch.randn(N, H // 2, W // 2, 2 * c, device="cuda").half(); skip = torch.randn(N, H, W, c, device="cuda").half()
stl = torch.stack((low.double().sum((1, 2)), (low.double() ** 2).sum((1, 2))), 2).contiguous()
sts = torch.stack((skip.double().sum((1, 2)), (skip.double() ** 2).sum((1, 2))), 2).contiguous()
ctw = torch.randn(2 * c, c, 2, 2, device="cuda") * 0.25; ctb = torch.randn(c, device="cuda") * 0.2
w = torch.randn(c, 2 * c, 3, 3, device="cuda") / 12
wp = ops.pack_conv3x3(w); ctp = ops.pack_convt2x2(ctw); wtc = ops.pack_conv3x3_tc(wp, dt)
g1 = torch.ones(2 * c, device="cuda"); b1 = torch.zeros(2 * c, device="cuda")
s0 = ops.make_src(low, 2 * c, xform=ops.DG_X_CONVT2, stats=stl, gamma=g1, beta=b1, groups=8, ct_w=ctp, ct_b=ctb, ct_cout=c,
                  ct_w_tc=ops.pack_convt2x2_tc(ctp, dt))
s1 = ops.make_src(skip, c, stats=sts, gamma=g1[:c], beta=b1[:c], groups=8)
comp = ops.pack_dec_composite(ctp, ctb, wp, dt)
out = torch.empty(N, H, W, c, device="cuda", dtype=torch.float16)
stt = torch.zeros(N, c, 2, device="cuda", dtype=torch.float64)
for name, path in (("composite", 2), ("round-1 fused", 2 | 1024)):
    for it in range(3):
        ops.conv3x3_fused([s0, s1], wp, c, N, H, W, dt, out=out, out_stats=stt, path=path, weight_tc=wtc, weight_comp=comp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(10):
        ops.conv3x3_fused([s0, s1], wp, c, N, H, W, dt, out=out, out_stats=stt, path=path, weight_tc=wtc, weight_comp=comp)
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 10, "ms")
