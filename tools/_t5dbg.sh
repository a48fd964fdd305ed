timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/r2_bench_full_a.txt 2>&1; echo "bench rc=$?"; tail -c 6000 gpurun_out/r2_bench_full_a.txt
