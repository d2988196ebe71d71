#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet_bf16.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_s2t.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2t.log
timeout 300 python scripts/profile_layers.py > gpurun_out/layers_s2t.log 2>&1; grep "down0\|total" gpurun_out/layers_s2t.log
