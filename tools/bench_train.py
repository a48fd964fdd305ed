#!/usr/bin/env python
"""Training-step timing (BASELINE.json configs[3]): forward + L1 + backward + clip 1.0 + AdamW, batch B x 1x512x512.

    python tools/bench_train.py [--batch 32] [--steps 5] [--storage fp32] [--graph]
    torchrun --nproc-per-node N tools/bench_train.py ...     (data parallel: one flat-gradient all-reduce per step)
Prints one JSON line: images/s over all ranks and ms/step (CUDA events, max over ranks).  (The CPU oracle's step is timed by the
tests, not here: only tests/, smoke() and bench.py's cpu_baseline leg may touch oracle/.)
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import image_enhancement_deglaring_b200 as dg  # noqa: E402
from image_enhancement_deglaring_b200.train import FusedAdamW  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32, help="per-GPU batch")
ap.add_argument("--hw", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--storage", default="fp32")
ap.add_argument("--graph", action="store_true", help="also time the step replayed from one CUDA graph (train.GraphedTrainStep)")
ap.add_argument("--cpu-batch", type=int, default=0, help="ignored (kept for old command lines)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
net = dg.LightweightUNet(storage=a.storage)
net.load_state_dict(sd, strict=True)
net = net.cuda().train()
opt = FusedAdamW(net.parameters(), lr=0.002362532125818593, weight_decay=6.753784966611083e-05, max_grad_norm=1.0, capturable=a.graph)
crit = dg.L1Loss()   # drop-in for nn.L1Loss (fused seed)
x = torch.rand(a.batch, 1, a.hw, a.hw, generator=torch.Generator().manual_seed(rank)).cuda()
t = torch.rand(a.batch, 1, a.hw, a.hw, generator=torch.Generator().manual_seed(100 + rank)).cuda()


def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(net(x), t)
    loss.backward()
    opt.step()
    return loss


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms) / a.steps
line = {"metric": "unet_deglare_train_images_per_sec", "value": world * a.batch / (ms * 1e-3), "unit": "images/s",
        "n_gpus": world, "ms_per_step": ms, "batch_per_gpu": a.batch, "storage": a.storage, "loss": float(loss.detach())}
if a.graph:
    from image_enhancement_deglaring_b200.train import GraphedTrainStep  # noqa: E402
    g = GraphedTrainStep(net, opt, crit, x.shape)
    for _ in range(a.warmup):
        g(x, t)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(a.steps):
        gl = g(x, t)
    e1.record()
    torch.cuda.synchronize()
    gms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
    gms = float(gms) / a.steps
    line["cuda_graph"] = {"ms_per_step": gms, "value": world * a.batch / (gms * 1e-3), "loss": float(gl)}
if rank == 0:
    print(json.dumps(line), flush=True)
# a captured graph holds NCCL work: drain and leave without the communicator teardown (it did not return within minutes after a
# graph capture with NCCL ops on 2 GPUs; the process exit releases everything)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
