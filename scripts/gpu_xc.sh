#!/bin/bash
# x-combined conv bring-up: parity first, then per-layer profile for each variant
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet_bf16.py -x -q > gpurun_out/txc.log 2>&1
echo "xc tests exit $?"; tail -8 gpurun_out/txc.log
for v in 1 3 4; do
  SQ_XC=$v timeout 300 python scripts/profile_layers.py > gpurun_out/layers_xc$v.log 2>&1
  echo "== SQ_XC=$v"; grep "down0\|down1\|up1\|up0\|total" gpurun_out/layers_xc$v.log
done
