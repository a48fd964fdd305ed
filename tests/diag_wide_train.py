"""Diagnostic (GPU): LightweightUNet(features_start=64) training step (BASELINE.json configs[4]) -- forward + backward time with the
tensor-core backward (path 0) and the CUDA-core backward (path 1) of the same 16-bit tier, and the gradient agreement between them.

    python tests/diag_wide_train.py [batch [H W]]
"""
import os
import re
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
    grads, per = {}, {}
    paths = tuple(int(v) for v in os.environ.get("DG_DIAG_PATHS", "0,1").split(","))
    for path in paths:
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
        per[path] = {k: p.grad.detach().double().clone() for k, p in net.named_parameters()}
        print(f"path {path} ({'tensor-core' if path == 0 else 'CUDA-core'} backward): forward {ev[0].elapsed_time(ev[1]):.2f} ms, "
              f"L1 + backward {ev[1].elapsed_time(ev[2]):.2f} ms for {B}x{H}x{W}  ({dg.count_parameters(net)} parameters)")
        del net
        torch.cuda.empty_cache()
    if os.environ.get("DG_DIAG_PROFILE"):
        # per-kernel device time of ONE tensor-core-backward step (CUPTI through torch.profiler), grouped by kernel name
        from torch.profiler import ProfilerActivity, profile
        net = dg.LightweightUNet(features_start=64, storage="fp16")
        net.load_state_dict(sd, strict=True)
        net = net.cuda().train()
        crit(net(x), t).backward()
        net.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            crit(net(x), t).backward()
            torch.cuda.synchronize()
        rows = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                m = re.search(r"(\w+_kernel)", e.name)
                tm = re.search(r"_kernel<(.{0,60})", e.name)
                name = (m.group(1) + (" <" + tm.group(1) if tm else "")) if m else e.name[:70]
                r = rows.setdefault(name, [0, 0.0, 0.0])
                r[0] += 1
                r[1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
                r[2] = max(r[2], e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total)
        tot = sum(r[1] for r in rows.values())
        print(f"one step, {tot / 1e3:.2f} ms of kernels:")
        for name, r in sorted(rows.items(), key=lambda kv: -kv[1][1])[:28]:
            print(f"  {r[1] / 1e3:9.3f} ms  {100 * r[1] / tot:5.1f} %  x{r[0]:<3d} max {r[2] / 1e3:8.3f} ms  {name}")
    if len(grads) < 2:
        return
    rel = float((grads[0] - grads[1]).norm() / grads[1].norm())
    print(f"gradient rel-L2 difference tensor-core vs CUDA-core backward: {rel:.3e}")
    rows = sorted(((float((per[0][k] - per[1][k]).norm() / per[1][k].norm().clamp_min(1e-300)), k) for k in per[0]), reverse=True)
    print("largest per-tensor differences: " + ", ".join(f"{k} {v:.2e}" for v, k in rows[:6]))


if __name__ == "__main__":
    main()
