for i in 1 2; do timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "t5" 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20; done
timeout 600 python -m pytest tests/test_gpu_forward.py -q -m gpu 2>&1 | grep -E "^FAILED|passed|failed|Error" | head -20
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2_bench_t5_d.txt 2>&1; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_t5_d.txt").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print({k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()})
PY
DG_BATCH_SPLIT=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2_bench_t5_d0.txt 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_t5_d0.txt").read().strip().splitlines()[-1])
print("nosplit", d["value"], d["ms_per_step"])
PY
DG_T5_TRACE=1 DG_BATCH_SPLIT=0 python tools/profile_forward.py --batch 64 --forwards 2 2>&1 | tail -140 > gpurun_out/r2_t5_trace.txt
