#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/mini_run.py"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_first_fused" -s 2 -c 1 \
    -o gpurun_out/prof_fused -f $CMD > gpurun_out/ncu6.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu6.log
