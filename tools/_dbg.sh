timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "t5" --tb=short 2>&1 | grep -E "^E   Assert|^FAILED|passed|failed" | head -20
python tools/wide_layers.py 4 2>&1 | tail -21
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --no-extras > gpurun_out/r2_bench_x.txt 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_x.txt").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
print({k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()})
PY
DG_T5_TRACE=1 python tools/t5_layer_trace.py same 64 64 512 512 4 2>&1 | grep -E "trace|item +[1-3] " | tail -4 | cut -c1-230
