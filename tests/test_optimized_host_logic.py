"""CPU test of OptimizedUNet's host orchestration (which per-op C-ABI call reads which tensor, channel window and order):
`ops.*` is swapped for the torch stand-ins of tests/opt_standins.py -- restatements of the semantics include/deglare.h documents --
and the module's own `_run` / `_backward` / autograd bridge must then reproduce the oracle's forward and all 76 gradients.
The CUDA kernels themselves are tested on the GPU (tests/test_gpu_optimized.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import torch_unet as tpo

sys.path.insert(0, os.path.dirname(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import opt_standins  # noqa: E402
from make_golden import det_state_dict  # noqa: E402

dg = pytest.importorskip("image_enhancement_deglaring_b200")


def _rand(shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def _net(golden, monkeypatch):
    from image_enhancement_deglaring_b200 import ops
    opt_standins.install(monkeypatch, ops)
    g = golden("opt_rand.npz")
    tmpl = {k: tuple(int(s) for s in sh.split(",")) for k, sh in zip(g["keys"], g["shapes"])}
    sd = {k: torch.from_numpy(v) for k, v in det_state_dict(tmpl, seed=1234).items()}
    net = dg.OptimizedUNet()
    net.load_state_dict(sd, strict=True)
    return net, sd


@pytest.mark.parametrize("shape,seed", [((2, 1, 32, 32), 7), ((1, 1, 16, 48), 11)])
def test_orchestration_reproduces_oracle_gradients(golden, monkeypatch, shape, seed):
    net, sd = _net(golden, monkeypatch)
    x = _rand(shape, seed)
    gy = torch.randn(*shape, generator=torch.Generator().manual_seed(9)) / x.numel()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = tpo.optimized_forward(x, params)
    ref = dict(zip(params, torch.autograd.grad((out * gy).sum(), list(params.values()))))
    keep = {}
    y = net._run(x, keep)
    assert float((y - out.detach()).abs().max()) <= 1e-4 * max(1.0, float(out.detach().abs().max()))
    flat = torch.zeros(sum(p.numel() for p in net.parameters()))
    net._backward(keep, gy, flat)
    off = 0
    for k, p in net.named_parameters():
        got = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
        r = ref[k]
        assert float((got - r).abs().max()) <= 1e-9 + 2e-4 * float(r.abs().max()), k


def test_autograd_bridge_returns_gradients_in_parameter_order(golden, monkeypatch):
    from image_enhancement_deglaring_b200.model_optimized import _OptimizedUNetFn
    net, sd = _net(golden, monkeypatch)
    x, t = _rand((1, 1, 32, 32), 3), _rand((1, 1, 32, 32), 4)
    r = tpo.train_step(sd, x, t, forward=tpo.optimized_forward, max_norm=0.0)
    y = _OptimizedUNetFn.apply(net, x, *net.parameters())
    loss = torch.nn.L1Loss()(y, t)
    loss.backward()
    assert abs(float(loss.detach()) - r["loss"]) <= 1e-5 * max(1.0, r["loss"])
    for k, p in net.named_parameters():
        ref = r["grads"][k]
        assert p.grad is not None and p.grad.shape == ref.shape, k
        assert float((p.grad - ref).abs().max()) <= 2e-5 + 1e-3 * float(ref.abs().max()), k
    with pytest.raises(RuntimeError, match="CUDA"):
        net(x)                       # the public forward still refuses CPU tensors: no CPU fallback in the product


# ---- data parallel: the mean all-reduce of the flat gradient sits at the END of the bridge's backward ------------------
def _opt_ddp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image_enhancement_deglaring_b200 import ops
    from image_enhancement_deglaring_b200.model_optimized import _OptimizedUNetFn
    for name in opt_standins.STANDINS:
        setattr(ops, name, getattr(opt_standins, name))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "opt_rand.npz"), allow_pickle=False)
    tmpl = {k: tuple(int(v) for v in sh.split(",")) for k, sh in zip(g["keys"], g["shapes"])}
    sd = {k: torch.from_numpy(v) for k, v in det_state_dict(tmpl, seed=1234).items()}
    net = dg.OptimizedUNet()
    net.load_state_dict(sd, strict=True)
    x, t = _rand((2, 1, 16, 32), 3), _rand((2, 1, 16, 32), 4)
    y = _OptimizedUNetFn.apply(net, x[rank:rank + 1], *net.parameters())      # one image per rank
    torch.nn.L1Loss()(y, t[rank:rank + 1]).backward()
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert torch.equal(gathered[0], gathered[1])          # every rank holds the averaged gradient after backward
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_bridge_averages_gradients_over_two_ranks(golden, tmp_path):
    out = str(tmp_path / "flat.pt")
    mp.spawn(_opt_ddp_worker, args=(2, 29500 + (os.getpid() % 2000) + 7, out), nprocs=2, join=True)
    got = torch.load(out)
    g = golden("opt_rand.npz")
    tmpl = {k: tuple(int(v) for v in sh.split(",")) for k, sh in zip(g["keys"], g["shapes"])}
    sd = {k: torch.from_numpy(v) for k, v in det_state_dict(tmpl, seed=1234).items()}
    x, t = _rand((2, 1, 16, 32), 3), _rand((2, 1, 16, 32), 4)
    # equal shards: the mean over ranks of the per-rank mean losses is the global-batch loss of optimized_train.py:223
    r = tpo.train_step(sd, x, t, forward=tpo.optimized_forward, max_norm=0.0)
    want = torch.cat([r["grads"][k].reshape(-1) for k in sd if k in dict(dg.OptimizedUNet().named_parameters())])
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-5 + 1e-3 * float(want.abs().max())
