timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_ddp.py -q -m gpu --tb=short 2>&1 | grep -E "^E  |^FAILED|passed|failed|Error" | head -30
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_train_a.txt 2>&1; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_train_a.txt").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["train_step"])
PY
