#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/mini_run.py"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"first_conv|ccl_" -s 9 -c 9 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu4.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu4.log; cat gpurun_out/plain.log | tail -2
