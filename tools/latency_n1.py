#!/usr/bin/env python
"""Single-image latency (BASELINE.json configs[0] shape: batch 1 x 1x512x512): device forward, ORT-shaped session.run from
host buffers, uint8 variant, and the same forward replayed from a CUDA graph."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg
from image_enhancement_deglaring_b200.session import InferenceSession
sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
for storage in ("fp16", "fp32"):
    net = dg.LightweightUNet(storage=storage); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    x = torch.rand(1, 1, 512, 512).cuda()
    with torch.no_grad():
        for _ in range(5): y = net(x)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(200): y = net(x)
        torch.cuda.synchronize()
        dev = (time.perf_counter() - t) / 200
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                yg = net(x)
        g.replay(); torch.cuda.synchronize()
        err = float((yg - y).abs().max())
        t = time.perf_counter()
        for _ in range(200): g.replay()
        torch.cuda.synchronize()
        gr = (time.perf_counter() - t) / 200
    sess = InferenceSession(net)
    xh = x.cpu().numpy()
    for _ in range(5): sess.run(["output"], {"input": xh})
    t = time.perf_counter()
    for _ in range(200): sess.run(["output"], {"input": xh})
    e2e = (time.perf_counter() - t) / 200
    u = (xh * 255).astype(np.uint8)
    for _ in range(5): sess.run_u8(u)
    t = time.perf_counter()
    for _ in range(200): sess.run_u8(u)
    e8 = (time.perf_counter() - t) / 200
    print(f"{storage}: device forward {dev*1e3:.3f} ms, graph replay {gr*1e3:.3f} ms (max diff {err:.1e}), session.run {e2e*1e3:.3f} ms, run_u8 {e8*1e3:.3f} ms")
