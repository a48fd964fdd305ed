#!/usr/bin/env python
"""BASELINE.json configs[4] status: LightweightUNet(features_start=64) ("sweep.py wider variant", 31.0 M parameters), forward at
batch x 1x512x512.  Channels 64 / 128 run on the tensor-core kernels, 256 / 512 / 1024 on the generic CUDA-core kernel."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for storage in ("fp16", "fp32"):
    torch.manual_seed(42)
    net = dg.LightweightUNet(features_start=64, storage=storage).cuda().eval()
    x = torch.rand(B, 1, 512, 512).cuda()
    with torch.no_grad():
        for _ in range(2): net(x)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(3): net(x)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    print(f"fs=64 {storage}: {dt*1e3:.1f} ms per batch {B} = {B/dt:.1f} img/s = {384.7e9*B/dt/1e12:.1f} TFLOP/s ({dg.count_parameters(net)} params)")
