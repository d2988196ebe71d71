#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet_bf16.py -x -q > gpurun_out/t3d.log 2>&1
echo "bf16 tests exit $?"; tail -3 gpurun_out/t3d.log
timeout 600 python scripts/bench_config4.py > gpurun_out/config4.log 2>&1
echo "c4 exit $?"; grep -v '^{' gpurun_out/config4.log | grep -v "down[2-4]\|up[2-3]\|maxpool"
timeout 300 python scripts/profile_layers.py > gpurun_out/layers.log 2>&1; tail -3 gpurun_out/layers.log
