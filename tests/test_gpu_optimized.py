"""GPU parity of the OptimizedUNet surface (src/optimized_model.py) against reference-generated golden vectors."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import torch_unet as tpo

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import det_state_dict  # noqa: E402

pytestmark = pytest.mark.gpu

dg = pytest.importorskip("image_enhancement_deglaring_b200")


def _rand(shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def _sd(golden):
    g = golden("opt_rand.npz")
    tmpl = {k: tuple(int(s) for s in sh.split(",")) for k, sh in zip(g["keys"], g["shapes"])}
    return g, {k: torch.from_numpy(v) for k, v in det_state_dict(tmpl, seed=1234).items()}


def test_optimized_matches_reference_golden(golden):
    g, sd = _sd(golden)
    net = dg.OptimizedUNet()
    assert list(net.state_dict().keys()) == list(g["keys"])
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    for shape, seed, key in (((2, 1, 64, 64), 3, "y_2x64x64_seed3"), ((1, 1, 48, 80), 4, "y_1x48x80_seed4")):
        with torch.no_grad():
            y = net(_rand(shape, seed).cuda()).cpu().numpy()
        ref = g[key]
        err = np.abs(y - ref).max()
        assert err <= 2e-4 * max(1.0, np.abs(ref).max()), (key, err)


def test_optimized_16bit_storage_close_to_oracle(golden):
    _, sd = _sd(golden)
    x = _rand((2, 1, 128, 128), 9)
    with torch.no_grad():
        ref = tpo.optimized_forward(x, sd).numpy()
    scale = max(1.0, float(np.abs(ref).max()))
    for storage, tol in (("fp16", 1e-2), ("bf16", 8e-2)):
        net = dg.OptimizedUNet(storage=storage)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().eval()
        with torch.no_grad():
            y = net(x.cuda()).cpu().numpy()
        assert np.abs(y - ref).max() <= tol * scale, storage


def test_optimized_errors(golden):
    _, sd = _sd(golden)
    net = dg.OptimizedUNet()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="multiples of 16"):
            net(torch.zeros(1, 1, 40, 40, device="cuda"))
        with pytest.raises(RuntimeError, match="CUDA"):
            net(torch.zeros(1, 1, 16, 16))
    with pytest.raises(NotImplementedError, match="input image"):
        net.train()(torch.zeros(1, 1, 16, 16, device="cuda", requires_grad=True))


# ---- training (src/optimized_model.py:118-158 under optimized_train.py:220-233) -------------------------------------------------
def _oracle_grads(sd, x, gy):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = tpo.optimized_forward(x, params)
    grads = torch.autograd.grad((out * gy).sum(), list(params.values()))
    return out.detach(), dict(zip(params.keys(), grads))


def _train_net(sd, **kw):
    net = dg.OptimizedUNet(**kw)
    net.load_state_dict(sd, strict=True)
    return net.cuda().train()


@pytest.mark.parametrize("shape,seed", [((2, 1, 64, 64), 7), ((1, 1, 48, 80), 11)])
def test_optimized_backward_matches_oracle_fp32(golden, shape, seed):
    """Every one of the 76 parameter gradients for an explicit output gradient (no sign() at the loss) against the autograd
    oracle, and for the golden shape against the REFERENCE module's own gradients (tests/golden/opt_train.npz)."""
    _, sd = _sd(golden)
    x = _rand(shape, seed)
    gy = torch.randn(*shape, generator=torch.Generator().manual_seed(9)) / x.numel()
    _, ref = _oracle_grads(sd, x, gy)
    net = _train_net(sd)
    net(x.cuda()).backward(gy.cuda())
    bad = []
    for k, p in net.named_parameters():
        assert p.grad is not None and p.grad.shape == ref[k].shape, k
        r = ref[k].numpy()
        err = float(np.abs(p.grad.cpu().numpy() - r).max())
        if err > 1e-10 + 3e-4 * float(np.abs(r).max()):
            bad.append((k, err, float(np.abs(r).max())))
    assert not bad, bad
    if shape == (2, 1, 64, 64):
        g = golden("opt_train.npz")
        for k, p in net.named_parameters():
            got = p.grad.cpu().numpy().reshape(-1)
            norm = float(g["gy_gnorm/" + k])
            assert abs(np.sqrt((got.astype(np.float64) ** 2).sum()) - norm) <= 3e-4 * norm + 1e-10, k
            assert np.abs(got[:32] - g["gy_ghead/" + k]).max() <= 1e-10 + 1e-3 * np.abs(g["gy_ghead/" + k]).max(), k


def test_optimized_reference_loop_step_fp32(golden):
    """optimized_train.py:220-233 verbatim on OptimizedUNet: zero_grad, forward, L1Loss, backward, clip_grad_norm_(1.0), AdamW step
    (FusedAdamW); the golden step of the reference module pins loss / total norm, the autograd oracle every gradient."""
    from image_enhancement_deglaring_b200.train import FusedAdamW
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import TRAIN_LR, TRAIN_WD
    g = golden("opt_train.npz")
    _, sd = _sd(golden)
    x, t = _rand((2, 1, 64, 64), 7), _rand((2, 1, 64, 64), 8)
    r = tpo.train_step(sd, x, t, forward=tpo.optimized_forward, lr=TRAIN_LR, weight_decay=TRAIN_WD, max_norm=1.0)
    net = _train_net(sd)
    opt = FusedAdamW(net.parameters(), lr=TRAIN_LR, weight_decay=TRAIN_WD)
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.L1Loss()(net(x.cuda()), t.cuda())
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * max(1.0, float(g["loss"]))
    grads = {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()}
    for k, gr in grads.items():
        ref = r["grads"][k].numpy()
        # sign(o - t) may flip on a handful of pixels where |o - t| ~ 1e-7: a few 1/numel steps of slack on top of the fp32 bound
        assert float(np.abs(gr.numpy() - ref).max()) <= 2e-5 + 1e-3 * float(np.abs(ref).max()), k
    total = torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
    opt.step()
    assert abs(float(total) - float(g["total_norm"])) <= 2e-3 * float(g["total_norm"])
    clipped, _ = tpo.clip_grad_norm(grads, 1.0)
    want = tpo.adamw_step({k: v.clone() for k, v in sd.items()}, clipped, {}, TRAIN_LR, TRAIN_WD)
    for k, p in net.named_parameters():
        assert float((p.detach().cpu() - want[k]).abs().max()) <= 1e-6, k
    # the packed-weight caches see the update
    with torch.no_grad():
        y2 = net.eval()(x.cuda()).cpu()
        ref2 = tpo.optimized_forward(x, {k: p.detach().cpu() for k, p in net.named_parameters()})
    assert float((y2 - ref2).abs().max()) <= 2e-4 * max(1.0, float(ref2.abs().max()))


def test_optimized_reference_amp_gradscaler_branch_verbatim(golden):
    """optimized_train.py:201-219 (the branch the reference takes on every CUDA device) with the reference's own torch.optim.AdamW:
    autocast + GradScaler.scale / unscale_ / step / update + clip_grad_norm_ run unchanged on the drop-in OptimizedUNet and land on
    the reference module's fp32 step (the module keeps its own precision under autocast; the power-of-two loss scale cancels)."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import TRAIN_LR, TRAIN_WD
    g = golden("opt_train.npz")
    _, sd = _sd(golden)
    net = _train_net(sd)
    opt = torch.optim.AdamW(net.parameters(), lr=TRAIN_LR, weight_decay=TRAIN_WD)
    scaler = torch.amp.GradScaler("cuda")
    crit = torch.nn.L1Loss()
    x, t = _rand((2, 1, 64, 64), 7).cuda(), _rand((2, 1, 64, 64), 8).cuda()
    opt.zero_grad(set_to_none=True)                       # :201
    with torch.amp.autocast("cuda"):                      # :205
        outputs = net(x)                                  # :206
        loss = crit(outputs, t)                           # :207
    scaler.scale(loss).backward()                         # :210
    scaler.unscale_(opt)                                  # :214
    total = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)   # :215
    scaler.step(opt)                                      # :218
    scaler.update()                                       # :219
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * max(1.0, float(g["loss"]))
    assert abs(float(total) - float(g["total_norm"])) <= 2e-3 * float(g["total_norm"])
    loose = n = 0
    for k, p in net.named_parameters():
        got = p.detach().cpu().numpy().reshape(-1)
        err = np.abs(got[:32] - g["newhead/" + k])
        assert err.max() <= 2 * TRAIN_LR, k
        loose += int((err > 2e-5).sum())
        n += err.size
        assert abs(float(got.astype(np.float64).sum()) - float(g["newsum/" + k])) <= 2e-5 * got.size ** 0.5 + 2 * TRAIN_LR, k
    assert loose <= 2e-3 * n, (loose, n)


@pytest.mark.parametrize("storage,global_tol", [("fp16", 1.5e-2), ("bf16", 6e-2)])
def test_optimized_backward_16bit_storage_is_close(golden, storage, global_tol):
    """16-bit storage of the saved activations; tensor-core weight / data gradients (bf16 operands, fp32 accumulate) where the
    channel pair is covered, the CUDA-core kernels elsewhere."""
    _, sd = _sd(golden)
    shape = (2, 1, 64, 64)
    x = _rand(shape, 7)
    gy = torch.randn(*shape, generator=torch.Generator().manual_seed(9)) / x.numel()
    _, ref = _oracle_grads(sd, x, gy)
    net = _train_net(sd, storage=storage)
    net(x.cuda()).backward(gy.cuda())
    num = den = 0.0
    worst = ("", 0.0)
    for k, p in net.named_parameters():
        d = float(((p.grad.cpu().double() - ref[k].double()) ** 2).sum())
        n = float((ref[k].double() ** 2).sum())
        num += d
        den += n
        if n > 0 and (d / n) ** 0.5 > worst[1]:
            worst = (k, (d / n) ** 0.5)
    assert (num / den) ** 0.5 <= global_tol, ((num / den) ** 0.5, worst)


def test_optimized_second_backward_accumulates_and_eval_still_works(golden):
    _, sd = _sd(golden)
    net = _train_net(sd)
    x = _rand((1, 1, 32, 32), 3).cuda()
    gy = torch.ones(1, 1, 32, 32, device="cuda") / 1024
    net(x).backward(gy)
    g1 = {k: p.grad.clone() for k, p in net.named_parameters()}
    net(x).backward(gy)
    for k, p in net.named_parameters():
        assert torch.allclose(p.grad, 2 * g1[k], rtol=1e-4, atol=1e-9), k
    with torch.no_grad():
        ref = tpo.optimized_forward(x.cpu(), sd)
        assert float((net.eval()(x).cpu() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))


# ---- the OptimizedUNet-only backward kernels against torch autograd on the same device -------------------------------------------
def test_grad_gather_matches_torch():
    from image_enhancement_deglaring_b200 import ops
    gen = torch.Generator().manual_seed(0)
    N, H, W, C = 2, 6, 10, 16
    a = torch.randn(N, H, W, 2 * C, generator=gen).cuda()
    sc = torch.rand(N, C, generator=gen).cuda()
    b = torch.randn(N, H // 2, W // 2, C, generator=gen).cuda()
    u = torch.randn(N, 2 * H, 2 * W, C + 4, generator=gen).cuda()
    add = torch.randn(N, C, generator=gen).cuda()
    got = ops.grad_gather(N, H, W, C, a=a, off_a=C, a_scale=sc, b=b, u=u, off_u=4, add=add)
    want = (a[..., C:] * sc[:, None, None, :]
            + 0.25 * b.repeat_interleave(2, 1).repeat_interleave(2, 2)
            + u[..., 4:].reshape(N, H, 2, W, 2, C).sum(dim=(2, 4))
            + add[:, None, None, :])
    assert float((got - want).abs().max()) <= 1e-5
    got = ops.grad_gather(N, H, W, C, u=u[..., :C].contiguous())
    assert float((got - u[..., :C].reshape(N, H, 2, W, 2, C).sum(dim=(2, 4))).abs().max()) <= 1e-5
    # scalar path: a channel count / window that is not a multiple of 4
    a3 = torch.randn(N, H, W, 7, generator=gen).cuda()
    got = ops.grad_gather(N, H, W, 3, a=a3, off_a=2)
    assert float((got - a3[..., 2:5]).abs().max()) == 0.0


@pytest.mark.parametrize("dtype", ["fp32", "fp16"])
def test_attention_backward_kernels_match_torch(dtype):
    """d(scale) and the ChannelAttention MLP backward against autograd of src/optimized_model.py:185-202 restated in torch."""
    from image_enhancement_deglaring_b200 import ops
    tdt = {"fp32": torch.float32, "fp16": torch.float16}[dtype]
    code = {"fp32": ops.DG_F32, "fp16": ops.DG_F16}[dtype]
    gen = torch.Generator().manual_seed(1)
    N, H, W, C, hid, groups = 2, 8, 12, 32, 8, 4
    raw = torch.randn(N, H, W, C, generator=gen).to(tdt).cuda()
    gamma = (1 + 0.2 * torch.randn(C, generator=gen)).cuda()
    beta = (0.2 * torch.randn(C, generator=gen)).cuda()
    w1 = (torch.randn(hid, C, generator=gen) / C ** 0.5).cuda().requires_grad_(True)
    w2 = (torch.randn(C, hid, generator=gen) / hid ** 0.5).cuda().requires_grad_(True)
    d = torch.randn(N, H, W, 2 * C, generator=gen).cuda()
    r = raw.float()
    stats = torch.stack((r.double().sum(dim=(1, 2)), (r.double() ** 2).sum(dim=(1, 2))), dim=-1).contiguous()   # [N,C,2]
    A = torch.nn.functional.silu(torch.nn.functional.group_norm(r.permute(0, 3, 1, 2), groups, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    A = A.detach().requires_grad_(True)
    mean = A.mean(dim=(1, 2))
    att = torch.sigmoid(torch.nn.functional.silu(mean @ w1.t()) @ w2.t())
    out = A * att[:, None, None, :]
    att.retain_grad()
    (out * d[..., C:]).sum().backward()
    dscale = ops.scale_bwd_sum(raw, stats, gamma, beta, groups, code, N, H, W, C, d, C)
    assert float((dscale.float() - att.grad).abs().max()) <= 2e-4 * float(att.grad.abs().max())
    act_sum = A.detach().double().sum(dim=(1, 2)).contiguous()
    dw1 = torch.zeros(hid, C, device="cuda")
    dw2 = torch.zeros(C, hid, device="cuda")
    add = ops.channel_attention_bwd(act_sum, float(H * W), w1.detach(), w2.detach(), dscale, dw1, dw2)
    assert float((dw1 - w1.grad).abs().max()) <= 3e-4 * float(w1.grad.abs().max())
    assert float((dw2 - w2.grad).abs().max()) <= 3e-4 * float(w2.grad.abs().max())
    # dL/dA = d * att (direct) + add (through the mean): autograd's A.grad holds both
    want_add = A.grad - d[..., C:] * att.detach()[:, None, None, :]
    assert float((add[:, None, None, :] - want_add).abs().max()) <= 3e-4 * float(want_add.abs().max()) + 1e-7


def test_optimized_cuda_graph_training_step_equals_eager_loop(golden):
    """train.GraphedTrainStep on OptimizedUNet: zero_grad, forward, L1, the per-op backward (~190 launches), clip, AdamW captured once
    and replayed; same losses and parameters as the eager loop, and an eager forward after the replays sees the updated weights."""
    from image_enhancement_deglaring_b200.train import FusedAdamW, GraphedTrainStep
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import TRAIN_LR, TRAIN_WD
    _, sd = _sd(golden)
    xs = [_rand((2, 1, 32, 48), 30 + i).cuda() for i in range(3)]
    ts = [_rand((2, 1, 32, 48), 40 + i).cuda() for i in range(3)]
    crit = torch.nn.L1Loss()
    # eager loop
    a = _train_net(sd)
    oa = FusedAdamW(a.parameters(), lr=TRAIN_LR, weight_decay=TRAIN_WD, max_grad_norm=1.0)
    la = []
    for x, t in zip(xs, ts):
        oa.zero_grad(set_to_none=True)
        loss = crit(a(x), t)
        loss.backward()
        oa.step()
        la.append(float(loss.detach()))
    # captured step
    b = _train_net(sd)
    ob = FusedAdamW(b.parameters(), lr=TRAIN_LR, weight_decay=TRAIN_WD, max_grad_norm=1.0, capturable=True)
    step = GraphedTrainStep(b, ob, crit, xs[0].shape)
    lb = [float(step(x, t)) for x, t in zip(xs, ts)]
    for u, v in zip(la, lb):
        assert abs(u - v) <= 2e-5 * max(1.0, abs(u)), (la, lb)
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        d = (pa.detach() - pb.detach()).abs()
        assert float(d.max()) <= 2 * TRAIN_LR, k             # Adam turns atomics-order noise on a near-zero gradient into up to lr
        moved = float((pa.detach().cpu() - sd[k]).norm())
        assert float(d.norm()) <= 0.05 * moved + 1e-6, k
    with torch.no_grad():
        ya, yb = a.eval()(xs[0]).cpu(), b.eval()(xs[0]).cpu()
    assert float((ya - yb).abs().max()) <= 5e-3 * max(1.0, float(ya.abs().max()))
