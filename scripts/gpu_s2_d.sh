#!/bin/bash
# session 2, call D: prep reductions (batched loads), full gpu suite, smoke, bench
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_aux.py > gpurun_out/aux_s2d.log 2>&1; grep -v '^{' gpurun_out/aux_s2d.log | tail -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s2d.json 2> gpurun_out/bench_s2d.err
echo "bench exit $?"; cat gpurun_out/bench_s2d.json; tail -3 gpurun_out/bench_s2d.err
