#!/bin/bash
# session 2, call B: two-phase W3 column pass, square augment tiles, HBM traffic-mix probe
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_weightmap.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_s2b.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_s2b.log
timeout 600 python scripts/bench_aux.py > gpurun_out/aux_s2b.log 2>&1; grep -v '^{' gpurun_out/aux_s2b.log | tail -12
timeout 300 python scripts/hbm_probe.py > gpurun_out/hbm_probe.log 2>&1; cat gpurun_out/hbm_probe.log
