"""Diagnostic (GPU): per-tensor gradient errors of OptimizedUNet's backward against the autograd oracle, every storage tier.

    python tests/diag_optimized_train.py [H W [N]]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import det_state_dict  # noqa: E402

import image_enhancement_deglaring_b200 as dg  # noqa: E402
from oracle import torch_unet as tpo  # noqa: E402  (checker only)


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    g = np.load(os.path.join(ROOT, "tests", "golden", "opt_rand.npz"))
    tmpl = {k: tuple(int(s) for s in sh.split(",")) for k, sh in zip(g["keys"], g["shapes"])}
    sd = {k: torch.from_numpy(v) for k, v in det_state_dict(tmpl, seed=1234).items()}
    x = torch.rand(N, 1, H, W, generator=torch.Generator().manual_seed(7))
    gy = torch.randn(N, 1, H, W, generator=torch.Generator().manual_seed(9)) / x.numel()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = tpo.optimized_forward(x, params)
    ref = dict(zip(params, torch.autograd.grad((out * gy).sum(), list(params.values()))))
    for storage in ("fp32", "fp16", "bf16"):
        net = dg.OptimizedUNet(storage=storage)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().train()
        y = net(x.cuda())
        y.backward(gy.cuda())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            net.zero_grad(set_to_none=True)
            net(x.cuda()).backward(gy.cuda())
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        num = den = 0.0
        rows = []
        for k, p in net.named_parameters():
            d = float(((p.grad.cpu().double() - ref[k].double()) ** 2).sum())
            n = float((ref[k].double() ** 2).sum())
            num += d
            den += n
            rows.append(((d / max(n, 1e-300)) ** 0.5, float((p.grad.cpu() - ref[k]).abs().max()), float(ref[k].abs().max()), k))
        rows.sort(reverse=True)
        print(f"{storage}: forward max-abs err {float((y.detach().cpu() - out.detach()).abs().max()):.3e} "
              f"(|y| max {float(out.detach().abs().max()):.3f}); gradient global rel-L2 {(num / den) ** 0.5:.3e}; "
              f"fwd+bwd {ms:.2f} ms for {N}x{H}x{W}")
        for r in rows[:8]:
            print(f"    rel-L2 {r[0]:.3e}  max-abs {r[1]:.3e} of {r[2]:.3e}  {r[3]}")


if __name__ == "__main__":
    main()
