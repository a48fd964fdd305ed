"""Diagnostic (GPU): OptimizedUNet training step (zero_grad, forward, L1, per-op backward, clip 1.0, AdamW) issued eagerly from Python vs
replayed from one CUDA graph (train.GraphedTrainStep), CUDA events, mean of 10.

    python tests/diag_optimized_graph.py [storage]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg  # noqa: E402
from image_enhancement_deglaring_b200 import _lib  # noqa: E402
from image_enhancement_deglaring_b200.train import FusedAdamW, GraphedTrainStep  # noqa: E402


def main():
    storage = sys.argv[1] if len(sys.argv) > 1 else "fp16"
    crit = torch.nn.L1Loss()
    for shape in ((2, 1, 64, 64), (4, 1, 256, 256), (8, 1, 512, 512)):
        torch.manual_seed(0)
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(1)).cuda()
        t = torch.rand(*shape, generator=torch.Generator().manual_seed(2)).cuda()
        out = []
        for graphed in (False, True):
            torch.manual_seed(0)        # the same initial weights for both flavours
            net = dg.OptimizedUNet(storage=storage).cuda().train()
            opt = FusedAdamW(net.parameters(), lr=2e-3, weight_decay=6e-5, max_grad_norm=1.0, capturable=graphed)
            if graphed:
                step = GraphedTrainStep(net, opt, crit, shape)
                run = lambda: step(x, t)
            else:
                def run():
                    opt.zero_grad(set_to_none=True)
                    loss = crit(net(x), t)
                    loss.backward()
                    opt.step()
                    return loss
            for _ in range(3):
                run()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            n0 = _lib.launch_count()
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(10):
                loss = run()
            ev[1].record()
            torch.cuda.synchronize()
            out.append((ev[0].elapsed_time(ev[1]) / 10, (_lib.launch_count() - n0) // 10, float(loss)))
            del net, opt
            torch.cuda.empty_cache()
        (e_ms, e_l, e_loss), (g_ms, _, g_loss) = out
        print(f"{storage} {shape[0]}x{shape[2]}x{shape[3]}: eager {e_ms:.2f} ms per step ({e_l} library launches), "
              f"CUDA graph {g_ms:.2f} ms  (loss {e_loss:.4f} / {g_loss:.4f})")


if __name__ == "__main__":
    main()
