#!/usr/bin/env python
"""Per-parameter gradient error of the CUDA backward vs the autograd oracle (GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg  # noqa: E402
from oracle import torch_unet as tpo  # noqa: E402

sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
shape = tuple(int(v) for v in (sys.argv[1:5] or (2, 1, 64, 64)))
x = torch.rand(*shape, generator=torch.Generator().manual_seed(0))
t = torch.rand(*shape, generator=torch.Generator().manual_seed(1))
r = tpo.train_step(sd, x, t, max_norm=0.0)
net = dg.LightweightUNet(path=int(os.environ.get("DG_PATH", "1")), storage=os.environ.get("DG_STORAGE", "fp32"))
net.load_state_dict(sd, strict=True)
net = net.cuda().train()
loss = torch.nn.L1Loss()(net(x.cuda()), t.cuda())
loss.backward()
print("loss", float(loss), r["loss"])
for k, p in net.named_parameters():
    ref = r["grads"][k].numpy()
    got = p.grad.cpu().numpy()
    err = np.abs(got - ref).max()
    print(f"{k:24s} max|ref| {np.abs(ref).max():.3e}  err {err:.3e}  rel {err / (np.abs(ref).max() + 1e-12):.2e}  "
          f"|got| {np.abs(got).max():.3e}")
