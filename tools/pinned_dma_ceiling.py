#!/usr/bin/env python
"""Ceiling of the fp32 host-I/O end-to-end path: concurrent pinned H2D + D2H copies on every rank with NO compute.

    python tools/pinned_dma_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pinned_dma_ceiling.py

Each rank moves the bytes `InferenceSession.run_pinned` moves for a batch of 64 fp32 512x512 images (64 MiB in, 64 MiB out, in 8-image
chunks on separate copy streams, as dg_lw_infer_host does), all ranks at once.  Prints per-rank and aggregate GB/s and the
images/s those bytes would allow -- the number bench.py's `e2e` (fp32 I/O) has to be read against."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    from image_enhancement_deglaring_b200.parallel import bind_to_gpu_numa_node
    bind_to_gpu_numa_node(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, H, W, chunk = 64, 512, 512, 8
hx = torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory(); hy = torch.empty_like(hx).pin_memory()
dx = torch.empty((B, 1, H, W), dtype=torch.float32, device="cuda"); dy = torch.empty_like(dx)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def one_batch():
    for i in range(0, B, chunk):
        with torch.cuda.stream(s_in):
            dx[i:i + chunk].copy_(hx[i:i + chunk], non_blocking=True)
        with torch.cuda.stream(s_out):
            hy[i:i + chunk].copy_(dy[i:i + chunk], non_blocking=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    one_batch()
barrier()
steps = 20
t = time.perf_counter()
for _ in range(steps):
    one_batch()
barrier()
dt = time.perf_counter() - t
tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
dt = float(tt.item())
bytes_per_rank = 2 * B * H * W * 4 * steps
if rank == 0:
    print(json.dumps({"what": "concurrent pinned H2D + D2H, no compute, all ranks at once", "n_gpus": world,
                      "GBps_per_rank_each_way": bytes_per_rank / 2 / dt / 1e9, "GBps_aggregate_both_ways": world * bytes_per_rank / dt / 1e9,
                      "images_per_s_ceiling_fp32_io": world * B * steps / dt, "images_per_s_ceiling_u8_io": 4 * world * B * steps / dt}))
if world > 1:
    dist.destroy_process_group()
