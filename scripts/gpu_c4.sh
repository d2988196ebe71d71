#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_unet_bf16.py tests/test_gpu_unet_fp32.py -x -q -k "3d" > gpurun_out/t3d.log 2>&1
echo "3d exit $?"; tail -3 gpurun_out/t3d.log
timeout 600 python scripts/bench_config4.py > gpurun_out/config4.log 2>&1
echo "c4 exit $?"; cat gpurun_out/config4.log | grep -v '^{'
