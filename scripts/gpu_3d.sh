#!/bin/bash
# UNet3D tensor-core bring-up: 3-D tests first (bounded), then the planar bf16 suite + bench
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet_bf16.py -x -q -k "unet3d or unsupported" > gpurun_out/t3d.log 2>&1
echo "3d exit $?" >> gpurun_out/t3d.log
tail -15 gpurun_out/t3d.log
timeout 900 python -m pytest tests/test_gpu_unet_bf16.py tests/test_gpu_unet_fp32.py -x -q > gpurun_out/t2d.log 2>&1
echo "2d exit $?" >> gpurun_out/t2d.log
tail -5 gpurun_out/t2d.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_3dchange.json 2> gpurun_out/bench_3dchange.err
tail -2 gpurun_out/bench_3dchange.json
