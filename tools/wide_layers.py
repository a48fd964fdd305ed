#!/usr/bin/env python
"""Per-kernel device times of the wide variant LightweightUNet(features_start=64) (BASELINE.json configs[4]) via dg_lw_profile."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg
from image_enhancement_deglaring_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(42)
net = dg.LightweightUNet(features_start=64, storage="fp16").cuda().eval()
x = torch.rand(B, 1, 512, 512).cuda()
with torch.no_grad():
    y = net(x)
ws = torch.empty(net.workspace_bytes(B, 512, 512), dtype=torch.uint8, device="cuda")
buf = (C.c_float * 19)()
acc = [0.0] * 19
for it in range(4):
    _lib.check(_lib.load().dg_lw_profile(C.byref(net.c_params()), x.data_ptr(), y.data_ptr(), B, 512, 512, ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream().cuda_stream, buf))
    if it:
        acc = [a + b for a, b in zip(acc, buf)]
names = ["enc1.0", "enc1.3", "enc2.0", "enc2.3", "enc3.0", "enc3.3", "enc4.0", "enc4.3", "bott.0", "bott.3", "up4+dec4.0", "dec4.3",
         "up3+dec3.0", "dec3.3", "up2+dec2.0", "dec2.3", "up1+dec1.0", "dec1.3", "head"]
f = [64 << i for i in range(5)]; px = [(512 >> i) ** 2 for i in range(5)]
mac = [px[0] * 9 * 1 * f[0], px[0] * 9 * f[0] * f[0]]
for l in range(1, 5):
    mac += [px[l] * 9 * f[l - 1] * f[l], px[l] * 9 * f[l] * f[l]]
for l in (3, 2, 1, 0):
    mac += [px[l + 1] * f[l + 1] * f[l] * 4 + px[l] * 9 * 2 * f[l] * f[l], px[l] * 9 * f[l] * f[l]]
mac += [px[0] * f[0]]
tot = 0
for n, a, m in zip(names, acc, mac):
    ms = a / 3
    tot += ms
    print(f"{n:12s} {ms:7.3f} ms  {2 * m * B / (ms * 1e-3) / 1e12:7.1f} TFLOP/s")
print(f"sum {tot:.3f} ms per batch {B}: {2 * sum(mac) * B / (tot * 1e-3) / 1e12:.1f} TFLOP/s")
