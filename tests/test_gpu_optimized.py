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
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 1, 16, 16, device="cuda"))
