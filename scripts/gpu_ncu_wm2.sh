#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
CMD="python scripts/wm_run.py"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:edt_|inst_|w1_table' -s 10 -c 5 -o gpurun_out/prof_wm2 -f $CMD > gpurun_out/ncu_wm2.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_wm2.log
