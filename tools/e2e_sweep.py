#!/usr/bin/env python
"""e2e (host buffers in/out) throughput of InferenceSession.run_pinned vs chunk size, plus raw pinned PCIe copy rates."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg
from image_enhancement_deglaring_b200.session import InferenceSession
sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
net = dg.LightweightUNet(storage="fp16"); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
B, H, W = 64, 512, 512
hx = torch.rand(B, 1, H, W).pin_memory(); hy = torch.empty(B, 1, H, W).pin_memory()
dx = torch.empty(B, 1, H, W, device="cuda")
for name, fn in (("H2D", lambda: dx.copy_(hx, non_blocking=True)), ("D2H", lambda: hy.copy_(dx, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print(f"{name}: {B*H*W*4/dt/1e9:.1f} GB/s ({dt*1e3:.2f} ms per 64 MB)")
with torch.no_grad():
    for nb in (8, 16, 32, 64):
        x = dx[:nb].contiguous(); net(x); torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10): net(x)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
        print(f"device forward batch {nb}: {dt*1e3:.3f} ms ({nb/dt:.0f} img/s)")
for chunk in (4, 8, 16, 32, 64):
    sess = InferenceSession(net, chunk=chunk)
    for _ in range(2): sess.run_pinned(hx, hy)
    t = time.perf_counter()
    for _ in range(10): sess.run_pinned(hx, hy)
    dt = (time.perf_counter() - t) / 10
    print(f"e2e chunk {chunk:2d}: {dt*1e3:.2f} ms/batch64  {B/dt:.0f} img/s")
# asynchronous submit / wait (dg_lw_infer_host_submit): `depth` batches in flight, fp32 and uint8 host buffers
for dtype in (torch.float32, torch.uint8):
    pairs = [((torch.rand(B, 1, H, W) * (255 if dtype == torch.uint8 else 1)).to(dtype).pin_memory(),
              torch.empty(B, 1, H, W, dtype=dtype).pin_memory()) for _ in range(3)]
    for chunk in (8, 16, 32, 64):
        for depth in (2, 3):
            sess = InferenceSession(net, chunk=chunk)
            pend = []
            def step(i):
                if len(pend) == depth: sess.wait(pend.pop(0))
                pend.append(sess.submit(*pairs[i % 3], chunk=chunk))
            def drain():
                while pend: sess.wait(pend.pop(0))
            for i in range(4): step(i)
            drain(); t = time.perf_counter()
            for i in range(12): step(i)
            drain(); dt = (time.perf_counter() - t) / 12
            print(f"submit/wait {str(dtype)[6:]:7s} chunk {chunk:2d} depth {depth}: {dt*1e3:.2f} ms/batch64  {B/dt:.0f} img/s")
            del sess
