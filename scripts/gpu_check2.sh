#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet_bf16.py -m gpu -q --no-header -rf --timeout 300 -x > gpurun_out/pytest_bf16.log 2>&1
echo "pytest bf16 exit $?"
tail -30 gpurun_out/pytest_bf16.log
timeout 300 python scripts/profile_layers.py > gpurun_out/layers.log 2>&1
echo "profile exit $?"
tail -45 gpurun_out/layers.log
