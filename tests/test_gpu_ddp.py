"""Data-parallel training over NCCL (2 GPUs).  The one collective -- a mean all-reduce of the flat gradient -- sits at the END OF
BACKWARD, so everything the reference loop does after `loss.backward()` (GradScaler.unscale_ + inf check, clip_grad_norm_,
optimizer.step; optimized_train.py:210-219 / :226-233) sees the global-batch gradient on every rank: a 2-rank run of the
reference's loop, VERBATIM, must reproduce the single-process step on the global batch (mean of per-rank L1 means == global
mean for equal shards).  Skipped with fewer than 2 GPUs (the gloo world-2 tests in test_host_logic.py cover the host logic)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LR, WD = 0.002362532125818593, 6.753784966611083e-05


def _make_opt(kind, net):
    from image_enhancement_deglaring_b200.train import FusedAdamW
    if kind == "fused_clip":
        return FusedAdamW(net.parameters(), lr=LR, weight_decay=WD, max_grad_norm=1.0)
    if kind == "fused":
        return FusedAdamW(net.parameters(), lr=LR, weight_decay=WD)
    return torch.optim.AdamW(net.parameters(), lr=LR, weight_decay=WD)


def _step(kind, net, opt, scaler, x, t):
    crit = torch.nn.L1Loss()
    opt.zero_grad(set_to_none=True)
    if kind == "fused_clip":                         # clip folded into the optimizer
        loss = crit(net(x), t)
        loss.backward()
        opt.step()
    elif scaler is None:                             # optimized_train.py:220-233 (fp32 branch) verbatim
        loss = crit(net(x), t)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
        opt.step()
    else:                                            # optimized_train.py:204-219 (CUDA branch) verbatim
        with torch.amp.autocast("cuda"):
            loss = crit(net(x), t)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
        scaler.step(opt)
        scaler.update()
    return float(loss.detach())


CASES = [("fused_clip", False), ("fused", False), ("fused", True), ("adamw", True)]


def _run_all(net_factory, x, t):
    out = {}
    for kind, amp in CASES:
        net = net_factory()
        opt = _make_opt(kind, net)
        scaler = torch.amp.GradScaler("cuda") if amp else None
        for _ in range(2):                           # two steps: the second one sees moments + updated weights
            _step(kind, net, opt, scaler, x, t)
        out[f"{kind}/{int(amp)}"] = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    return out


def _worker(rank, world, port, out):
    import image_enhancement_deglaring_b200 as dg
    from image_enhancement_deglaring_b200.parallel import shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))

    def factory():
        net = dg.LightweightUNet(path=1)
        net.load_state_dict(sd, strict=True)
        return net.cuda().train()

    x = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(0))
    t = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(1))
    lo, hi = shard_range(4, rank, world)
    res = _run_all(factory, x[lo:hi].cuda(), t[lo:hi].cuda())
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_reference_loop_equals_global_batch_step(tmp_path, best_sd):
    import image_enhancement_deglaring_b200 as dg
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, out), nprocs=2, join=True)
    got = torch.load(out)

    def factory():
        net = dg.LightweightUNet(path=1)
        net.load_state_dict(best_sd, strict=True)
        return net.cuda().train()

    x = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(0)).cuda()
    t = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(1)).cuda()
    want = _run_all(factory, x, t)
    for case in want:
        loose = total = 0
        for k, v in want[case].items():
            err = np.abs(v.numpy() - got[case][k].numpy())
            assert err.max() <= 4 * LR, (case, k)
            loose += int((err > 4e-5).sum())
            total += err.size
        assert loose <= 2e-4 * total, (case, loose, total)


def test_single_process_reference_loops_agree(best_sd):
    """The four optimiser / loop flavours of the 2-rank test are the same step on one GPU (this part runs on the 1-GPU box)."""
    import image_enhancement_deglaring_b200 as dg

    def factory():
        net = dg.LightweightUNet(path=1)
        net.load_state_dict(best_sd, strict=True)
        return net.cuda().train()

    x = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(0)).cuda()
    t = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(1)).cuda()
    res = _run_all(factory, x, t)
    base = res["fused_clip/0"]
    for case, sd in res.items():
        loose = total = 0
        for k, v in sd.items():
            err = np.abs(v.numpy() - base[k].numpy())
            assert err.max() <= 4 * LR, (case, k)
            loose += int((err > 4e-5).sum())
            total += err.size
        assert loose <= 2e-4 * total, (case, loose, total)
