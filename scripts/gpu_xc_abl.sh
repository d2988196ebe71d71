#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for d in 0 1 2 3 4; do
  SQ_XC=2 SQ_XC_DBG=$d timeout 300 python scripts/profile_layers.py > gpurun_out/layers_abl$d.log 2>&1
  echo "== SQ_XC=2 DBG=$d"; grep "down0/conv2\|down1\|up1/conv\|up0/conv\|total" gpurun_out/layers_abl$d.log
done
