#!/usr/bin/env python
"""profiles/ncu_traffic.json (per-layer DRAM bytes of one batch-64 forward, read by bench.py for `roofline.traffic`) from the launch
list `tools/ncu_summary.py --json` writes for `ncu --set full -k regex:"^(conv|dec8|head|convt)" python tools/profile_forward.py
--batch 64 --forwards 1`.  The forward runs as two half-batch launches per layer (half 0 first, then half 1, 21 kernels each: the 19
layers plus the two stand-alone ConvTranspose kernels of levels 4 and 3, which are counted with the conv they feed).

    python tools/ncu_traffic.py gpurun_out/launches.json > profiles/ncu_traffic.json
"""
import json
import sys

LAYERS = ["enc1.0", "enc1.3", "enc2.0", "enc2.3", "enc3.0", "enc3.3", "enc4.0", "enc4.3", "bottleneck.0", "bottleneck.3", "up4+dec4.0",
          "up4+dec4.0", "dec4.3", "up3+dec3.0", "up3+dec3.0", "dec3.3", "up2+dec2.0", "dec2.3", "up1+dec1.0", "dec1.3", "head"]
rows = json.load(open(sys.argv[1]))
if len(rows) % len(LAYERS):
    sys.exit(f"{len(rows)} launches is not a multiple of {len(LAYERS)} kernels per half-batch forward")
out = {}
for i, r in enumerate(rows):
    d = out.setdefault(LAYERS[i % len(LAYERS)], {"read": 0.0, "write": 0.0, "traffic": 0.0, "us_serialised": 0.0, "kernels": []})
    d["read"] += r["read"]
    d["write"] += r["write"]
    d["traffic"] += r["read"] + r["write"]
    d["us_serialised"] += r["us"]
    name = r["kernel"].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "")
    if name not in d["kernels"]:
        d["kernels"].append(name)
json.dump(out, sys.stdout, indent=1)
