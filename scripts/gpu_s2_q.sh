#!/bin/bash
# session 2, call Q: ncu --set full of the two DPX column passes
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cols_dpx" -s 2 -c 2 -o gpurun_out/prof_dpx -f python scripts/aux_run.py > gpurun_out/ncu_dpx.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_dpx.log
