#!/bin/bash
# session 2, call A: augment kernel + first-conv restructure
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_augment.py tests/test_gpu_unet_bf16.py tests/test_gpu_unet_fp32.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_s2a.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_s2a.log
timeout 300 python scripts/profile_layers.py > gpurun_out/layers_s2a.log 2>&1; cat gpurun_out/layers_s2a.log
timeout 600 python scripts/bench_aux.py > gpurun_out/aux_s2a.log 2>&1; grep -v '^{' gpurun_out/aux_s2a.log | tail -12
