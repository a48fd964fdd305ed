#!/usr/bin/env python
"""Per-layer share of a forward from an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv ...
python bench.py --steps 2 --warmup 1 --no-train --no-extras --no-cpu-baseline`) beside the live CUDA-event times of a bench line.

    python tools/ncu_launch_shares.py gpurun_out/launches.csv gpurun_out/bench.json > profiles/rNN_ncu_launches.txt

A batch-64 forward is two half-batch launches per layer; a half-forward = 21 kernels (19 layers, the two deep ConvTransposes are
their own kernels and are added to their decoder layer).  ncu times are serialised and cold-cache: the SHARES must agree, not the
absolute times."""
import csv
import json
import sys

LAYERS = ["enc1.0", "enc1.3", "enc2.0", "enc2.3", "enc3.0", "enc3.3", "enc4.0", "enc4.3", "bottleneck.0", "bottleneck.3", "up4+dec4.0",
          "dec4.3", "up3+dec3.0", "dec3.3", "up2+dec2.0", "dec2.3", "up1+dec1.0", "dec1.3", "head"]
SLOTS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 10, 11, 12, 12, 13, 14, 15, 16, 17, 18]   # kernel k of a half-forward -> layer
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    name = r["Kernel Name"].replace("void ", "").replace("dg::<unnamed>::", "").replace("dg::", "").split("<")[0].split("(")[0]
    rows.append((name, float(r["Metric Value"].replace(",", "")) * (1e-3 if r["Metric Unit"] in ("ns", "nsecond") else 1.0)))
starts = [i for i, (n, _) in enumerate(rows) if n == "conv_first_tc_kernel"]
halves = [rows[i:i + 21] for i in starts if i + 21 <= len(rows) and rows[i + 20][0].startswith("head")]
acc, names = [0.0] * 19, [set() for _ in range(19)]
for h in halves:
    for k, (n, us) in enumerate(h):
        acc[SLOTS[k]] += us / len(halves)
        names[SLOTS[k]].add(n)
bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
live = bench["roofline"]["per_kernel"]
tot, ltot = sum(acc), sum(live[l]["ms"] for l in LAYERS)
print(f"# {len(halves)} half-batch forwards of 21 kernels in the first {len(rows)} launches; bench line: {bench['value']:.0f} img/s, {bench['ms_per_step']:.4f} ms per step")
print("# layer          ncu us per half-forward   share    live ms (batch 64)   share")
for l, a in zip(LAYERS, acc):
    print(f"{l:14s} {a:10.1f} {a / tot:20.3f} {live[l]['ms']:14.4f} {live[l]['ms'] / ltot:14.3f}")
print(f"{'sum':14s} {tot:10.1f} {1.0:20.3f} {ltot:14.4f} {1.0:14.3f}")
print("# kernel names per layer: " + ", ".join(f"{l}={'+'.join(sorted(n))}" for l, n in zip(LAYERS, names)))
