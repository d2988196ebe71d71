#!/bin/bash
mkdir -p gpurun_out
export N=4
CMD="python scripts/bench_aux.py"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"edt_|inst_" -s 8 -c 6 \
    -o gpurun_out/prof_wm -f $CMD > gpurun_out/ncu5.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu5.log
