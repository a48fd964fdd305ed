#!/usr/bin/env python
"""Instruction hot spots of one profiled launch: SASS grouped by execution count, with opcode mix and stall samples.
    python tools/ncu_hot.py <rep> <launch index>"""
import collections, csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:110])
hdr = rows[1]
iA, iE, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
seq = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: e = float(r[iE].replace(",", ""))
    except ValueError: continue
    seq.append((e, float(r[iS].replace(",", "") or 0), r[iA]))
seq = seq[:len(seq) // 2] if len(seq) > 1 and seq[0][2] == seq[len(seq) // 2][2] else seq
tot = sum(e for e, _, _ in seq); ts = sum(s for _, s, _ in seq) or 1
def op(src):
    t = src.split(); o = t[1] if t[0].startswith("@") else t[0]; return o.split(".")[0]
groups = collections.OrderedDict()
for e, s, src in seq:
    g = groups.setdefault(e, [0, 0, collections.Counter()]); g[0] += e; g[1] += s; g[2][op(src)] += 1
print(f"{len(seq)} SASS instrs, {tot/1e6:.1f} M warp-inst executed")
for e, g in sorted(groups.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"exec={e:10.0f} n={sum(g[2].values()):4d} inst%={g[0]/tot*100:5.1f} samp%={g[1]/ts*100:5.1f} " + " ".join(f"{k}:{v}" for k, v in g[2].most_common(14)))
ops = collections.Counter(); smp = collections.Counter()
for e, s, src in seq: ops[op(src)] += e; smp[op(src)] += s
print("opcode mix: " + " ".join(f"{k}:{v/tot*100:.1f}%({smp[k]/ts*100:.0f}%s)" for k, v in ops.most_common(18)))
