#!/bin/bash
# session 2, call M: level-0 fusion -- equality with the two-launch path, tests, per-layer times
cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/fuse_check.py > gpurun_out/fuse_check.log 2>&1
echo "fuse_check exit $?"; tail -8 gpurun_out/fuse_check.log
timeout 900 python -m pytest tests/test_gpu_unet_bf16.py -m gpu -q --no-header -rf --timeout 600 -x > gpurun_out/pytest_s2m.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_s2m.log
SQ_FUSE_FIRST=0 timeout 300 python scripts/profile_layers.py > gpurun_out/layers_s2m_off.log 2>&1; grep "down0\|total" gpurun_out/layers_s2m_off.log
timeout 300 python scripts/profile_layers.py > gpurun_out/layers_s2m_on.log 2>&1; grep "down0\|total" gpurun_out/layers_s2m_on.log
