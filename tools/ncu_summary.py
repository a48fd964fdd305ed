#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into one line per kernel launch: time, DRAM bytes, pipe utilisation, stalls.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--names a,b,c] [--json launches.json] > profiles/xxx.txt

--json also writes [{"kernel", "us", "read", "write"}] per launch (DRAM bytes), the input of profiles/ncu_traffic.json.  Run it on the
GPU box right after the capture and delete the .ncu-rep there: a --set full report of one forward is larger than what gpurun copies back.
"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
names = None
if "--names" in sys.argv:
    names = sys.argv[sys.argv.index("--names") + 1].split(",")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


def tounit(r, k, base):
    v = f(r, k)
    u = units[ix[k]] if k in ix else ""
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1.0)
    return v * scale / base


stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print(f"{'kernel':14s} {'us':>8s} {'rdMB':>7s} {'wrMB':>7s} {'dram%':>6s} {'sm%':>5s} {'issue%':>6s} {'tens%':>6s} {'xu%':>5s} "
      f"{'warps%':>6s} {'regs':>4s} {'Minst':>7s}  top stalls")
for i, r in enumerate(data):
    nm = names[i] if names and i < len(names) else r[ix["Kernel Name"]][:14]
    st = sorted(((f(r, s), s.split("stalled_")[1].split("_per_")[0]) for s in stalls), reverse=True)[:4]
    print(f"{nm:14s} {tounit(r, 'gpu__time_duration.sum', 1e-6):8.1f} {tounit(r, 'dram__bytes_read.sum', 1e6):7.1f} "
          f"{tounit(r, 'dram__bytes_write.sum', 1e6):7.1f} {f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{f(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {r[ix['launch__registers_per_thread']]:>4s} "
          f"{f(r, 'smsp__inst_executed.sum') / 1e6:7.1f}  " + " ".join(f"{n}={v:.1f}" for v, n in st))

if "--json" in sys.argv:
    with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
        json.dump([{"kernel": r[ix["Kernel Name"]], "us": tounit(r, "gpu__time_duration.sum", 1e-6),
                    "read": tounit(r, "dram__bytes_read.sum", 1.0), "write": tounit(r, "dram__bytes_write.sum", 1.0)} for r in data], fh, indent=1)
