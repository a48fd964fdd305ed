#!/bin/bash
# quick per-kernel view of one bench run (used between kernel experiments)
out=${1:-gpurun_out/bench_k.txt}
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out 2>&1
python - "$out" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "e2e_u8", round(d.get("e2e_u8", {}).get("value", 0)))
print({k: v["ms"] for k, v in d["roofline"]["per_kernel"].items()})
PY
