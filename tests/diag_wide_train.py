"""Diagnostic (GPU): LightweightUNet(features_start=64) training step (BASELINE.json configs[4]) -- forward + backward time with the
tensor-core backward (path 0) and the CUDA-core backward (path 1) of the same 16-bit tier, and the gradient agreement between them.

    python tests/diag_wide_train.py [batch [H W]]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    torch.manual_seed(0)
    base = dg.LightweightUNet(features_start=64, storage="fp16")
    sd = {k: v.clone() for k, v in base.state_dict().items()}
    x = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(1)).cuda()
    t = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(2)).cuda()
    crit = torch.nn.L1Loss()
    grads = {}
    for path in (0, 1):
        net = dg.LightweightUNet(features_start=64, storage="fp16", path=path)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().train()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for it in range(3):
            net.zero_grad(set_to_none=True)
            if it == 2:
                ev[0].record()
            y = net(x)
            if it == 2:
                ev[1].record()
            crit(y, t).backward()
            if it == 2:
                ev[2].record()
        torch.cuda.synchronize()
        grads[path] = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double()
        print(f"path {path} ({'tensor-core' if path == 0 else 'CUDA-core'} backward): forward {ev[0].elapsed_time(ev[1]):.2f} ms, "
              f"L1 + backward {ev[1].elapsed_time(ev[2]):.2f} ms for {B}x{H}x{W}  ({dg.count_parameters(net)} parameters)")
        del net
        torch.cuda.empty_cache()
    rel = float((grads[0] - grads[1]).norm() / grads[1].norm())
    print(f"gradient rel-L2 difference tensor-core vs CUDA-core backward: {rel:.3e}")


if __name__ == "__main__":
    main()
