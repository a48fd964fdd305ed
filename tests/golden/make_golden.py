#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REFERENCE itself.

Runs only in the build container (needs /root/reference, which is absent on the GPU
box); the resulting .npz files are committed and are what pins `oracle/` (and through
it the CUDA path) to the reference.  Nothing here is imported by tests at run time
except `det_state_dict`, the deterministic weight generator shared with the tests.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

TRAIN_LR = 0.002362532125818593      # optimized_train.py:42
TRAIN_WD = 6.753784966611083e-05     # optimized_train.py:52


def det_state_dict(template, seed):
    """Deterministic (numpy legacy RandomState) weights for a {key: shape} template.

    conv / linear weights ~ N(0, 1/fan_in) * 1.4, GN gamma ~ 1 + 0.2 N, GN/conv bias ~ 0.2 N.
    Used for architectures that ship no checkpoint (OptimizedUNet, wide LightweightUNet)."""
    rs = np.random.RandomState(seed)
    out = {}
    for k, shape in template.items():
        shape = tuple(shape)
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:])) if len(shape) == 4 else shape[1]
            if "upconv" in k and len(shape) == 4 and shape[2] == 2:
                fan_in = shape[0]  # ConvTranspose2d weight is [Cin, Cout, 2, 2]
            out[k] = (rs.standard_normal(shape) * (1.4 / np.sqrt(fan_in))).astype(np.float32)
        elif k.endswith(".weight"):
            out[k] = (1.0 + 0.2 * rs.standard_normal(shape)).astype(np.float32)
        else:
            out[k] = (0.2 * rs.standard_normal(shape)).astype(np.float32)
    return out


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    import torch
    from PIL import Image

    torch.set_num_threads(os.cpu_count())
    ref_model = _load(os.path.join(REF, "src/model.py"), "ref_model")
    ref_opt = _load(os.path.join(REF, "src/optimized_model.py"), "ref_optimized_model")
    sd = torch.load(os.path.join(ROOT, "weights/best_model.pth"))

    net = ref_model.LightweightUNet()
    net.load_state_dict(sd, strict=True)
    net.eval()

    # ---- 1. the two real sample images through the /infer preprocessing (api/app.py:136-157)
    out = {}
    for i, name in enumerate(("test_input1.png", "test_input2.png"), 1):
        img = np.array(Image.open(os.path.join(REF, "api", name)))
        gray = np.array(Image.fromarray(img).convert("L"))
        gray = np.array(Image.fromarray(gray).resize((512, 512), Image.LANCZOS))
        x = torch.from_numpy(gray.astype(np.float32) / 255.0)[None, None]
        with torch.no_grad():
            y = net(x)
        out[f"x{i}_u8"] = gray
        out[f"y{i}"] = y[0, 0].numpy()
        print(name, float(y.min()), float(y.max()), float(y.mean()))
    np.savez_compressed(os.path.join(HERE, "lw_png.npz"), **out)

    # ---- 2. seeded random inputs, with every raw conv output (forward hooks)
    out = {}
    taps = {}
    hooks = []
    for mname, mod in net.named_modules():
        if isinstance(mod, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
            hooks.append(mod.register_forward_hook(
                lambda m, i, o, key=mname: taps.__setitem__(key, o.detach().numpy().copy())))
    x = torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        y = net(x)
    out["y_2x64x64_seed0"] = y.numpy()
    for k, v in taps.items():
        out["tap/" + k] = v
    for h in hooks:
        h.remove()
    x = torch.rand(1, 1, 96, 80, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        out["y_1x96x80_seed1"] = net(x).numpy()
    x = torch.rand(3, 1, 16, 16, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        out["y_3x16x16_seed2"] = net(x).numpy()
    x = torch.rand(2, 1, 512, 512, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        y = net(x)
    out["y_2x512x512_seed0_row255"] = y[:, 0, 255, :].numpy()
    out["y_2x512x512_seed0_stats"] = np.array(
        [float(y.double().sum()), float((y.double() ** 2).sum()), float(y.min()), float(y.max())])
    np.savez_compressed(os.path.join(HERE, "lw_rand.npz"), **out)

    # ---- 3. one fp32 training step through the reference's own call sequence
    #         (optimized_train.py:220-233 with L1Loss :439, AdamW :440-446, clip 1.0 :230)
    out = {}
    net = ref_model.LightweightUNet()
    net.load_state_dict(sd, strict=True)
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=TRAIN_LR, weight_decay=TRAIN_WD)
    crit = torch.nn.L1Loss()
    x = torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(0))
    t = torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(1))
    opt.zero_grad(set_to_none=True)
    loss = crit(net(x), t)
    loss.backward()
    for k, p in net.named_parameters():
        out["grad/" + k] = p.grad.numpy().copy()
    total = torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
    opt.step()
    out["loss"] = np.float32(loss.item())
    out["total_norm"] = np.float32(float(total))
    for k, p in net.named_parameters():
        out["new/" + k] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "lw_train.npz"), **out)
    print("train loss", loss.item(), "norm", float(total))

    # ---- 4. OptimizedUNet (no checkpoint ships) on deterministic weights
    out = {}
    net = ref_opt.OptimizedUNet()
    tmpl = {k: v.shape for k, v in net.state_dict().items()}
    osd = det_state_dict(tmpl, seed=1234)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in osd.items()}, strict=True)
    net.eval()
    x = torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        out["y_2x64x64_seed3"] = net(x).numpy()
    x = torch.rand(1, 1, 48, 80, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        out["y_1x48x80_seed4"] = net(x).numpy()
    out["n_params"] = np.int64(sum(p.numel() for p in net.parameters()))
    out["keys"] = np.array(list(tmpl.keys()))
    out["shapes"] = np.array([",".join(map(str, s)) for s in tmpl.values()])
    out["wsum"] = np.float64(sum(float(np.abs(v).astype(np.float64).sum()) for v in osd.values()))
    np.savez_compressed(os.path.join(HERE, "opt_rand.npz"), **out)

    # ---- 5. wider LightweightUNet variants (constructor arg features_start)
    out = {}
    for fs, hw in ((16, 64), (64, 32)):
        net = ref_model.LightweightUNet(features_start=fs)
        tmpl = {k: v.shape for k, v in net.state_dict().items()}
        wsd = det_state_dict(tmpl, seed=100 + fs)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in wsd.items()}, strict=True)
        net.eval()
        x = torch.rand(2, 1, hw, hw, generator=torch.Generator().manual_seed(5))
        with torch.no_grad():
            out[f"y_fs{fs}_2x{hw}x{hw}_seed5"] = net(x).numpy()
        out[f"keys_fs{fs}"] = np.array(list(tmpl.keys()))
        out[f"shapes_fs{fs}"] = np.array([",".join(map(str, s)) for s in tmpl.values()])
    # odd widths exercise the group-divisor search (src/model.py:71-86)
    net = ref_model.LightweightUNet(in_channels=3, out_channels=2, num_groups=8, features_start=12)
    tmpl = {k: v.shape for k, v in net.state_dict().items()}
    wsd = det_state_dict(tmpl, seed=112)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in wsd.items()}, strict=True)
    net.eval()
    x = torch.rand(2, 3, 32, 48, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        out["y_fs12_in3_out2_2x32x48_seed6"] = net(x).numpy()
    out["keys_fs12"] = np.array(list(tmpl.keys()))
    out["shapes_fs12"] = np.array([",".join(map(str, s)) for s in tmpl.values()])
    out["gn_groups_fs12"] = np.array([m.num_groups for m in net.modules() if isinstance(m, torch.nn.GroupNorm)])
    np.savez_compressed(os.path.join(HERE, "lw_variants.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    sys.exit(main())
