#!/bin/bash
for cfg in "2048 2048 4" "256 2048 4" "256 2048 2" "128 2048 4" "512 2048 2"; do
  set -- $cfg
  echo "=== H=$1 W=$2 N=$3"
  H=$1 W=$2 N=$3 timeout 300 python scripts/profile_layers.py 2>&1 | grep -E "down0|up0|total"
done
