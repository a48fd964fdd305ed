"""Diagnostic (not a test): which configuration breaks CUDA-graph capture of the training step.  python tests/diag_graph_capture.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, torch
sys.path.insert(0, %r)
import image_enhancement_deglaring_b200 as dg
from image_enhancement_deglaring_b200.train import FusedAdamW, GraphedTrainStep, L1Loss
hw, b, path, pre = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sd = torch.load(os.path.join(%r, "weights", "best_model.pth"))
net = dg.LightweightUNet(storage="fp16", path=path); net.load_state_dict(sd, strict=True); net = net.cuda().train()
opt = FusedAdamW(net.parameters(), lr=1e-3, max_grad_norm=1.0, capturable=True)
crit = L1Loss()
x = torch.rand(b, 1, hw, hw).cuda(); t = torch.rand(b, 1, hw, hw).cuda()
keep = os.environ.get("KEEP_LOSS") == "1"
for _ in range(pre):
    opt.zero_grad(set_to_none=True); loss = crit(net(x), t); loss.backward(); opt.step()
    if not keep: del loss
g = GraphedTrainStep(net, opt, crit, x.shape)
print("OK", float(g(x, t)))
''' % (ROOT, ROOT)
for name, env, args in [("hw512 b4 pre 2", {}, (512, 4, 0, 2)), ("hw512 b4 pre 12", {}, (512, 4, 0, 12)),
                        ("hw512 b4 pre 2 keep loss", {"KEEP_LOSS": "1"}, (512, 4, 0, 2)), ("hw64 b4 pre 2 keep loss", {"KEEP_LOSS": "1"}, (64, 4, 0, 2)),
                        ("hw64 b4 pre 12", {}, (64, 4, 0, 12))]:
    r = subprocess.run([sys.executable, "-c", CODE] + [str(a) for a in args], capture_output=True, text=True, env={**os.environ, **env}, timeout=120)
    err = [l for l in r.stderr.splitlines() if "Error" in l or "error" in l]
    print(f"{name:24s} rc={r.returncode} {r.stdout.strip()[-60:]} {err[-1][:160] if err and r.returncode else ''}", flush=True)
