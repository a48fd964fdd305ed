#!/usr/bin/env python
"""Short, fixed workload for ncu: `--forwards` LightweightUNet forwards at batch x 1 x hw x hw (fp16 storage by default).

    python tools/profile_forward.py --batch 64 --forwards 2
    ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 17 -c 17 -o gpurun_out/prof \
        python tools/profile_forward.py --batch 64 --forwards 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--hw", type=int, default=512)
ap.add_argument("--forwards", type=int, default=2)
ap.add_argument("--storage", default="fp16")
ap.add_argument("--path", type=int, default=0)
a = ap.parse_args()
net = dg.LightweightUNet(storage=a.storage, path=a.path)
net.load_state_dict(torch.load(os.path.join(ROOT, "weights", "best_model.pth")), strict=True)
net = net.cuda().eval()
x = torch.rand(a.batch, 1, a.hw, a.hw, generator=torch.Generator().manual_seed(0)).cuda()
with torch.no_grad():
    for _ in range(a.forwards):
        y = net(x)
torch.cuda.synchronize()
print("ok", float(y.mean()))
