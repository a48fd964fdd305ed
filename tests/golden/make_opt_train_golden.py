#!/usr/bin/env python
"""Golden vectors of ONE OptimizedUNet training step, produced by the REFERENCE module itself
(src/optimized_model.py under the fp32 call sequence of optimized_train.py:220-233: L1Loss :439, backward, clip 1.0 :230,
AdamW :440-446).  Build container only (needs /root/reference); the committed tests/golden/opt_train.npz pins
`oracle.torch_unet.train_step(forward=optimized_forward)` -- and through it the CUDA backward -- to the reference.

The model has 1.9 M parameters, so the fixture keeps every gradient's L2 norm and first 32 entries, the full gradient of
every tensor of at most 4096 elements (all GroupNorm affines, the attention MLPs, the head, the first conv), the loss, the
total norm and the same summary of the updated parameters.

    python tests/golden/make_opt_train_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, TRAIN_LR, TRAIN_WD, _load, det_state_dict  # noqa: E402

FULL_BELOW = 4096
HEAD = 32


def main():
    import torch

    torch.set_num_threads(os.cpu_count())
    ref_opt = _load(os.path.join(REF, "src/optimized_model.py"), "ref_optimized_model")
    net = ref_opt.OptimizedUNet()
    tmpl = {k: v.shape for k, v in net.state_dict().items()}
    osd = det_state_dict(tmpl, seed=1234)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in osd.items()}, strict=True)
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=TRAIN_LR, weight_decay=TRAIN_WD)
    crit = torch.nn.L1Loss()
    x = torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(7))
    t = torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(8))
    opt.zero_grad(set_to_none=True)
    y = net(x)
    loss = crit(y, t)
    loss.backward()
    out = {"loss": np.float32(loss.item()), "y": y.detach().numpy().copy()}
    for k, p in net.named_parameters():
        g = p.grad.numpy().reshape(-1)
        out["gnorm/" + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out["ghead/" + k] = g[:HEAD].copy()
        if g.size <= FULL_BELOW:
            out["grad/" + k] = p.grad.numpy().copy()
    total = torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
    opt.step()
    out["total_norm"] = np.float32(float(total))
    for k, p in net.named_parameters():
        v = p.detach().numpy().reshape(-1)
        out["newhead/" + k] = v[:HEAD].copy()
        out["newsum/" + k] = np.float64(v.astype(np.float64).sum())
    # a second fixture with an explicit output gradient (no sign() at the loss: insensitive to 1e-7 forward differences)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in osd.items()}, strict=True)
    net.zero_grad(set_to_none=True)
    gy = torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(9)) / (2 * 64 * 64)
    net(x).backward(gy)
    for k, p in net.named_parameters():
        g = p.grad.numpy().reshape(-1)
        out["gy_gnorm/" + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out["gy_ghead/" + k] = g[:HEAD].copy()
    np.savez_compressed(os.path.join(HERE, "opt_train.npz"), **out)
    print("loss", loss.item(), "norm", float(total), os.path.getsize(os.path.join(HERE, "opt_train.npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
