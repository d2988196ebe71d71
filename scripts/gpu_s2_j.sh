#!/bin/bash
# session 2, call J: final check -- full gpu suite, smoke, default bench, launch list under ncu
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -6 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_s2j.json 2> gpurun_out/bench_s2j.err
echo "bench exit $?"; cat gpurun_out/bench_s2j.json; tail -3 gpurun_out/bench_s2j.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s2.csv \
    python bench.py --steps 2 --warmup 1 --batch 2 --no-cpu-baseline > gpurun_out/ncu_launches_s2.log 2>&1
echo "ncu exit $?"
