"""Data-parallel training over NCCL (2 GPUs): one flat-gradient all-reduce per step must reproduce the single-process step
on the global batch (mean of per-rank L1 means == global mean for equal shards).  Skipped with fewer than 2 GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LR, WD = 0.002362532125818593, 6.753784966611083e-05


def _step(net, opt, x, t):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.L1Loss()(net(x), t)
    loss.backward()
    opt.step()
    return float(loss.detach())


def _worker(rank, world, port, out):
    import image_enhancement_deglaring_b200 as dg
    from image_enhancement_deglaring_b200.parallel import shard_range
    from image_enhancement_deglaring_b200.train import FusedAdamW
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
    net = dg.LightweightUNet(path=1)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    opt = FusedAdamW(net.parameters(), lr=LR, weight_decay=WD, max_grad_norm=1.0)
    x = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(0))
    t = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(1))
    lo, hi = shard_range(4, rank, world)
    _step(net, opt, x[lo:hi].cuda(), t[lo:hi].cuda())
    if rank == 0:
        torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_global_batch_step(tmp_path, best_sd):
    import image_enhancement_deglaring_b200 as dg
    from image_enhancement_deglaring_b200.train import FusedAdamW
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, out), nprocs=2, join=True)
    got = torch.load(out)
    net = dg.LightweightUNet(path=1)
    net.load_state_dict(best_sd, strict=True)
    net = net.cuda().train()
    opt = FusedAdamW(net.parameters(), lr=LR, weight_decay=WD, max_grad_norm=1.0)
    x = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(0)).cuda()
    t = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(1)).cuda()
    _step(net, opt, x, t)
    loose = total = 0
    for k, v in net.state_dict().items():
        err = np.abs(v.detach().cpu().numpy() - got[k].numpy())
        assert err.max() <= 2 * LR, k
        loose += int((err > 2e-5).sum())
        total += err.size
    assert loose <= 1e-4 * total, (loose, total)
