#!/bin/bash
# session 2, call L: label-and-localise tests and timing
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_label.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_s2l.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_s2l.log
timeout 600 python scripts/bench_aux.py > gpurun_out/aux_s2l.log 2>&1; grep "label_centroids" gpurun_out/aux_s2l.log | grep -v '^{'
