#!/usr/bin/env python
"""BASELINE.json configs[2], definition B: ONE 4096x4096 image through the network exactly (whole-image GroupNorm), rows sharded
over the ranks (whole_image.py: per conv ONE all-gather carrying the GroupNorm partial sums and the 2-row halos).

    python tools/whole_image_bench.py [--size 4096] [--storage fp16]                 # 1 GPU: the single-call forward AND 1 band
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/whole_image_bench.py

Times with CUDA events on every rank (after warm-up, barrier + synchronize on both sides), max over ranks; prints ONE JSON line
on rank 0: ms per image for the sharded path, for the un-sharded single call (1 GPU only), the 64-tile definition-A time for
scale, and the max-abs difference of the sharded result to the single-call result computed on rank 0."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--storage", default="fp16")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

import image_enhancement_deglaring_b200 as dg   # noqa: E402
from image_enhancement_deglaring_b200.tiling import infer_tiled   # noqa: E402
from image_enhancement_deglaring_b200.whole_image import BandGraph, DistComm, KernelBackend, band_with_halo, forward_band   # noqa: E402

sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
net = dg.LightweightUNet(storage=args.storage)
net.load_state_dict(sd, strict=True)
net = net.cuda().eval()
S = args.size
img = torch.rand(S, S, generator=torch.Generator().manual_seed(91)).cuda()
comm, be = DistComm(), KernelBackend(net)
band = band_with_halo(img, rank, world).contiguous()


def timed(fn):
    for _ in range(args.warmup):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


with torch.no_grad():
    ms_sharded = timed(lambda: forward_band(net, band, comm, S, be))
    y = forward_band(net, band, comm, S, be)
    full = comm.gather_rows(y.permute(1, 0, 2).contiguous()).permute(1, 0, 2)
    line = {"what": f"one {S}x{S} image, exact whole-image GroupNorm, rows sharded over ranks (SURVEY 8e definition B)",
            "n_gpus": world, "storage": args.storage, "ms_per_image_sharded": ms_sharded,
            "exchanges_per_image": {"all_gathers": 18, "packet": "C x 2 doubles of GroupNorm partial sums + 2 top rows + 2 bottom rows"}}
    try:   # the same band forward replayed from a CUDA graph (kernels + NCCL in one graph)
        g = BandGraph(net, comm, S, S, be)
        line["ms_per_image_sharded_cuda_graph"] = timed(lambda: g(band))
        yg = g(band)
        line["graph_equals_eager"] = bool(torch.equal(yg, y))
    except Exception as e:   # noqa: BLE001 -- report, the eager number stands
        line["cuda_graph_error"] = repr(e)[:300]
    if rank == 0:
        whole = net(img[None, None])[0]
        line["max_abs_sharded_vs_single_call"] = float((full - whole).abs().max())
    if world == 1:
        line["ms_per_image_single_call"] = timed(lambda: net(img[None, None]))
        line["ms_per_image_as_independent_512_tiles_definition_A"] = timed(lambda: infer_tiled(net, img, tile=512, batch=64))
if rank == 0:
    print(json.dumps(line), flush=True)
# a captured graph holds NCCL work: drop it and drain the device before tearing the communicator down, and do not let a slow
# NCCL teardown keep the box (measured: destroy_process_group after graph capture did not return within minutes on 2 GPUs)
g = None
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
