#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_weightmap.py -m gpu -q --no-header -rf --timeout 300 > gpurun_out/pytest_wm.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_wm.log
timeout 300 python scripts/bench_aux.py 2>&1 | tee gpurun_out/bench_aux.log | head -6
