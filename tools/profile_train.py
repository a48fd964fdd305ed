#!/usr/bin/env python
"""Short, fixed workload for ncu: `--steps` training steps (forward + L1 + backward + FusedAdamW) at batch x 1 x hw x hw.

    ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 17 -c 17 -o gpurun_out/wg \
        python tools/profile_train.py --batch 32 --steps 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg  # noqa: E402
from image_enhancement_deglaring_b200.train import FusedAdamW  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--hw", type=int, default=512)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--storage", default="fp16")
a = ap.parse_args()
net = dg.LightweightUNet(storage=a.storage)
net.load_state_dict(torch.load(os.path.join(ROOT, "weights", "best_model.pth")), strict=True)
net = net.cuda().train()
opt = FusedAdamW(net.parameters(), lr=2e-3, weight_decay=6e-5, max_grad_norm=1.0)
x = torch.rand(a.batch, 1, a.hw, a.hw, generator=torch.Generator().manual_seed(0)).cuda()
t = torch.rand(a.batch, 1, a.hw, a.hw, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(a.steps):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.L1Loss()(net(x), t)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("ok", float(loss))
