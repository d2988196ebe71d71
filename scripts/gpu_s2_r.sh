#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_weightmap.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_s2r.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2r.log
timeout 600 python scripts/bench_aux.py > gpurun_out/aux_s2r.log 2>&1; grep -v '^{' gpurun_out/aux_s2r.log | grep "weightmap"
