#!/usr/bin/env python
"""Per-kernel time of one training step (torch profiler, CUDA activities) to rank the backward kernels."""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg
from image_enhancement_deglaring_b200.train import FusedAdamW
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
storage = sys.argv[2] if len(sys.argv) > 2 else "fp32"
sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
net = dg.LightweightUNet(storage=storage); net.load_state_dict(sd, strict=True); net = net.cuda().train()
opt = FusedAdamW(net.parameters(), lr=2e-3, weight_decay=6e-5, max_grad_norm=1.0)
x = torch.rand(B, 1, 512, 512).cuda(); t = torch.rand(B, 1, 512, 512).cuda()
def step():
    opt.zero_grad(); loss = torch.nn.L1Loss()(net(x), t); loss.backward(); opt.step()
step(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name.replace("void dg::", "").replace("(anonymous namespace)::", "")
        if "conv3x3_generic_kernel" in k:
            k = "conv3x3_generic " + ("WGRAD" if "true" in k.split(">(")[0].split(",")[-1] or "(bool)1" in k else "fwd/dgrad") + " " + k.split("<")[1].split(",")[0]
        else:
            k = k.split("<")[0]
        k = k[:56]
        agg[k][0] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        agg[k][1] += 1
wg = [(e.name, (e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total)) for e in prof.events()
      if e.device_type == torch.autograd.DeviceType.CUDA and "conv3x3_generic_kernel" in e.name]
print("generic conv launches in order (ms):", " ".join(("W" if ("true" in n.split(">(")[0].split(",")[-1] or "(bool)1" in n) else "f") + f"{t/1e3:.2f}" for n, t in wg))
for pat in ("wgrad_tc_kernel", "dgrad_tc_kernel", "act_bwd_vec", "convt_bwd_data"):
    seq = [(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total) for e in prof.events()
           if e.device_type == torch.autograd.DeviceType.CUDA and pat in e.name and "convt_wgrad" not in e.name]
    print(pat, "in launch order (us):", " ".join(f"{t:.0f}" for t in seq))
tot = sum(v[0] for v in agg.values())
print(f"batch {B} storage {storage}: total device time {tot/1e3:.2f} ms")
for k, (t_us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {k:50s} {t_us/1e3:8.2f} ms  x{n:3d}  {t_us/tot*100:5.1f}%")
