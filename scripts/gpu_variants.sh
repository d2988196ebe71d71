#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 3; do
  echo "=== variant $v"
  SQ_TC_VARIANT=$v timeout 300 python scripts/profile_layers.py 2>&1 | grep -E "down0|down1|up1/conv|up0/conv|total"
done
