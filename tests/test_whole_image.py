"""CPU tests of the exact row-sharded whole-image path (whole_image.py; SURVEY 8e definition B): the halo bookkeeping, the
GroupNorm partial-sum all-reduce and the neighbour exchange, with a torch-functional stand-in for the CUDA kernels (test
infrastructure -- the product backend is the C-ABI and is exercised by tests/test_gpu_whole_image.py), single process with
thread bands and two processes over gloo, against the oracle applied to the WHOLE image."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import torch_unet as tpo   # noqa: E402  (checker)

_BLOCKS = ("enc1", "enc2", "enc3", "enc4", "bottleneck", "dec4", "dec3", "dec2", "dec1")


class TorchBackend:
    """Same interface as whole_image.KernelBackend, fp32 on CPU.  Fresh tensors are NaN-filled so that a halo row used before
    it was exchanged poisons the output."""

    def __init__(self, sd):
        self.sd = sd
        self.device = torch.device("cpu")

    def alloc(self, rows, W, channels):
        return torch.full((rows, W, channels), float("nan"))

    def _act(self, kind, t, coef, i):
        if kind == "image":
            return t[None, None]
        y = F.silu(t * coef[:, 0] + coef[:, 1]).permute(2, 0, 1)[None]
        if kind == "pool":
            return F.avg_pool2d(y, 2, 2)
        if kind == "convt":
            lvl = 9 - i // 2   # block dec4 (b = 5) consumes upconv4
            return F.conv_transpose2d(y, self.sd[f"upconv{lvl}.weight"], self.sd[f"upconv{lvl}.bias"], stride=2)
        return y

    def conv(self, i, srcs, out):
        x = torch.cat([self._act(kind, t, coef, i) for kind, t, coef, _ in srcs], 1)
        raw = F.conv2d(x, self.sd[f"{_BLOCKS[i // 2]}.{0 if i % 2 == 0 else 3}.weight"], None, 1, 1)[0]
        out.copy_(raw.permute(1, 2, 0))
        d = raw.double()
        return torch.stack((d.sum((1, 2)), (d * d).sum((1, 2))), -1)

    def head(self, t, coef, out):
        out.copy_(F.conv2d(self._act("same", t, coef, 17), self.sd["output_conv.weight"], self.sd["output_conv.bias"])[0])
        return out


def _net_and_sd(fs=8):
    import image_enhancement_deglaring_b200 as dg
    torch.manual_seed(3)
    net = dg.LightweightUNet(features_start=fs)
    with torch.no_grad():
        for p in net.parameters():   # GroupNorm affines away from (1, 0) so that a wrong (a, b) shows
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.2)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net, sd


@pytest.mark.parametrize("bands,H,W", [(1, 64, 32), (2, 64, 48), (4, 128, 32), (3, 192, 16)])
def test_band_algorithm_equals_whole_image_oracle(bands, H, W):
    from image_enhancement_deglaring_b200.whole_image import infer_whole_local
    net, sd = _net_and_sd()
    img = torch.rand(H, W, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        want = tpo.lightweight_forward(img[None, None], sd)[0]
        got = infer_whole_local(net, img, bands, backend=TorchBackend(sd))
    assert got.shape == want.shape
    assert not torch.isnan(got).any()
    assert float((got - want).abs().max()) <= 2e-5
    if bands > 1:   # and it is NOT the per-band (tiled) function: GroupNorm and the receptive field couple the bands
        with torch.no_grad():
            hb = H // bands
            tiled = torch.cat([tpo.lightweight_forward(img[None, None, r * hb:(r + 1) * hb], sd)[0] for r in range(bands)], 1)
        assert float((tiled - want).abs().max()) > 1e-3


def test_band_algorithm_other_width_and_groups():
    """features_start = 12: GroupNorm groups come from the divisor rule (6 groups of 2 ... ), channels per group > 1 at level 1."""
    from image_enhancement_deglaring_b200.whole_image import infer_whole_local
    net, sd = _net_and_sd(fs=12)
    img = torch.rand(64, 32, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        want = tpo.lightweight_forward(img[None, None], sd)[0]
        got = infer_whole_local(net, img, 2, backend=TorchBackend(sd))
    assert float((got - want).abs().max()) <= 2e-5


def test_band_rows_and_errors():
    from image_enhancement_deglaring_b200.whole_image import band_rows, band_with_halo, infer_whole_local
    assert band_rows(4096, 3, 8) == (1536, 2048)
    img = torch.zeros(128, 16)
    assert band_with_halo(img, 0, 4).shape == (34, 16) and band_with_halo(img, 1, 4).shape == (36, 16)
    assert band_with_halo(img, 3, 4).shape == (34, 16)
    with pytest.raises(RuntimeError):
        band_rows(96, 0, 2)      # 48 rows per band: not a multiple of 32
    net, sd = _net_and_sd()
    with pytest.raises(RuntimeError):   # a failing band must surface its error, not dead-lock the others
        infer_whole_local(net, torch.zeros(64, 24), 2, backend=TorchBackend(sd))
    with pytest.raises(RuntimeError):   # the product backend needs CUDA
        infer_whole_local(net, torch.zeros(64, 32), 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from image_enhancement_deglaring_b200.whole_image import infer_whole_sharded
    net, sd = _net_and_sd()
    img = torch.rand(64, 48, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        full = infer_whole_sharded(net, img, backend=TorchBackend(sd))
        band = infer_whole_sharded(net, img, gather=False, backend=TorchBackend(sd))
    torch.save({"full": full, "band": band}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_whole_image_sharded_world2_gloo(tmp_path):
    port = 33500 + (os.getpid() % 2000)
    out = str(tmp_path / "whole.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    _, sd = _net_and_sd()
    img = torch.rand(64, 48, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        want = tpo.lightweight_forward(img[None, None], sd)[0]
    assert torch.equal(r0["full"], r1["full"])
    assert float((r0["full"] - want).abs().max()) <= 2e-5
    assert torch.equal(r0["band"], r0["full"][:, :32]) and torch.equal(r1["band"], r0["full"][:, 32:])
