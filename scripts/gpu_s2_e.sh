#!/bin/bash
# session 2, call E: DPX column pass of W1, 64-frame e2e calls
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_weightmap.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_s2e.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_s2e.log
timeout 600 python scripts/bench_aux.py > gpurun_out/aux_s2e.log 2>&1; grep -v '^{' gpurun_out/aux_s2e.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2e.json 2> gpurun_out/bench_s2e.err
echo "bench exit $?"; cat gpurun_out/bench_s2e.json; tail -3 gpurun_out/bench_s2e.err
