"""GPU parity of the exact row-sharded whole-image path (whole_image.py; SURVEY 8e definition B, BASELINE.json configs[2]):
the per-op C-ABI on band-sized tensors with ready-made GroupNorm affines, partial-sum all-reduce and halo exchange per conv --
against the oracle applied to the WHOLE image and against the single-call whole-image forward.  Bands as threads on one GPU
(always) and as two NCCL ranks (when the box has two GPUs)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import torch_unet as tpo
from dg_testutil import psnr

pytestmark = pytest.mark.gpu

dg = pytest.importorskip("image_enhancement_deglaring_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net(sd, **kw):
    net = dg.LightweightUNet(**kw)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval()


@pytest.mark.parametrize("storage,tol", [("fp32", 2e-4), ("fp16", 5e-3)])
@pytest.mark.parametrize("bands,H,W", [(1, 64, 64), (2, 128, 96), (4, 256, 256), (3, 96, 272)])
def test_bands_on_one_gpu_match_whole_image_oracle(best_sd, storage, tol, bands, H, W):
    from image_enhancement_deglaring_b200.whole_image import infer_whole_local
    img = torch.rand(H, W, generator=torch.Generator().manual_seed(11 + bands))
    with torch.no_grad():
        want = tpo.lightweight_forward(img[None, None], best_sd)[0]
        net = _net(best_sd, storage=storage)
        got = infer_whole_local(net, img.cuda(), bands).cpu()
        whole = net(img[None, None].cuda())[0].cpu()
    err = float((got - want).abs().max())
    assert err <= tol, f"{storage} {bands} bands: max-abs vs oracle {err:.3e}"
    # the same function as the single whole-image call (statistics differ only by the order of the double sums)
    assert float((got - whole).abs().max()) <= (2e-5 if storage == "fp32" else 4e-3)
    if bands > 1:
        hb = H // bands
        with torch.no_grad():
            tiled = torch.cat([net(img[None, None, r * hb:(r + 1) * hb].cuda())[0].cpu() for r in range(bands)], 1)
        assert float((tiled - want).abs().max()) > 10 * tol      # per-band (definition A) is a different function


def test_wide_variant_bands(golden):
    """features_start = 64: every conv on the tcgen05 kernel with ready-made affines, stand-alone ConvTranspose at all levels."""
    from image_enhancement_deglaring_b200.whole_image import infer_whole_local
    torch.manual_seed(42)
    net = dg.LightweightUNet(features_start=64, storage="fp16").cuda().eval()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    img = torch.rand(64, 128, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = tpo.lightweight_forward(img[None, None], sd)[0]
        got = infer_whole_local(net, img.cuda(), 2).cpu()
    assert float((got - want).abs().max()) <= 5e-3


def test_4096_square_image_in_8_bands_fp16(best_sd):
    """configs[2] at full size: 8 bands of 512 rows (the 8-GPU partition) on one GPU vs the oracle on the whole image (north_star's
    16-bit bound) and vs the single-call whole-image forward (two fp16-storage evaluations of the same function through
    different kernels -- ready-made affines route the level-1 decoder to the un-composited kernel -- so the bound is looser)."""
    from image_enhancement_deglaring_b200.whole_image import infer_whole_local
    img = torch.rand(4096, 4096, generator=torch.Generator().manual_seed(91))
    torch.set_num_threads(os.cpu_count() or 1)
    net = _net(best_sd, storage="fp16")
    with torch.no_grad():
        want = tpo.lightweight_forward(img[None, None], best_sd)[0]
        whole = net(img[None, None].cuda())[0].cpu()
        got = infer_whole_local(net, img.cuda(), 8).cpu()
    err = float((got - want).abs().max())
    assert err <= 5e-3, f"8 bands vs oracle: {err:.3e}"
    assert psnr(got.numpy(), want.numpy()) >= 50.0
    assert float((got - whole).abs().max()) <= 8e-3
    assert psnr(got.numpy(), whole.numpy()) >= 60.0


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from image_enhancement_deglaring_b200.whole_image import infer_whole_sharded
    sd = torch.load(os.path.join(ROOT, "weights", "best_model.pth"))
    res = {}
    img = torch.rand(512, 384, generator=torch.Generator().manual_seed(21))
    for storage in ("fp32", "fp16"):
        net = dg.LightweightUNet(storage=storage)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().eval()
        with torch.no_grad():
            res[storage] = infer_whole_sharded(net, img.cuda()).cpu()
            res[storage + "/band"] = infer_whole_sharded(net, img.cuda(), gather=False).cpu()
    torch.cuda.synchronize()
    torch.save(res, f"{out}.{rank}")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gloo world-2 test in test_whole_image.py covers the host logic)")
def test_two_ranks_nccl(best_sd, tmp_path):
    port = 35500 + (os.getpid() % 2000)
    out = str(tmp_path / "whole.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    img = torch.rand(512, 384, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        want = tpo.lightweight_forward(img[None, None], best_sd)[0]
    for storage, tol in (("fp32", 2e-4), ("fp16", 5e-3)):
        assert torch.equal(r0[storage], r1[storage])
        assert float((r0[storage] - want).abs().max()) <= tol
        assert torch.equal(r0[storage + "/band"], r0[storage][:, :256]) and torch.equal(r1[storage + "/band"], r0[storage][:, 256:])
