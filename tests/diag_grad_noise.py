#!/usr/bin/env python
"""Run-to-run noise of the parameter gradients (fp32 atomics accumulate them in a different order every run)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg
sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
for storage in ("fp16", "fp32"):
    net = dg.LightweightUNet(storage=storage); net.load_state_dict(sd, strict=True); net = net.cuda().train()
    x = torch.rand(4, 1, 512, 512, generator=torch.Generator().manual_seed(5)).cuda()
    t = torch.rand(4, 1, 512, 512, generator=torch.Generator().manual_seed(6)).cuda()
    runs = []
    for _ in range(8):
        net.zero_grad(set_to_none=True)
        torch.nn.L1Loss()(net(x), t).backward()
        runs.append({k: p.grad.detach().clone() for k, p in net.named_parameters()})
    total = float(torch.sqrt(sum((v.double() ** 2).sum() for v in runs[0].values())))
    worst = []
    for k in runs[0]:
        dev = max(float((r[k] - runs[0][k]).norm()) for r in runs[1:])
        worst.append((dev / (float(runs[0][k].norm()) + 1e-30), dev / total, k))
    worst.sort(reverse=True)
    print(storage, "largest run-to-run deviation (relative to the tensor norm, relative to the whole gradient):")
    for a, b, k in worst[:5]: print(f"   {k:22s} {a:.2e}  {b:.2e}")
