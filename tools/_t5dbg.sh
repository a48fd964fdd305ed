timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "composite" --tb=short 2>&1 | grep -E "^E  |^FAILED|passed|failed|Error" | head -30
timeout 600 python -m pytest tests/test_gpu_forward.py -q -m gpu 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20
python tools/dec_layer_time.py
for P in 0 8; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --no-extras --path $P > gpurun_out/r2_bench_dec_$P.txt 2>&1; echo "bench path $P rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_dec_$P.txt").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print({k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()})
PY
done
python tools/accuracy_paths.py --paths 0,8
