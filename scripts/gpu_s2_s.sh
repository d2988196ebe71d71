#!/bin/bash
# session 2, call S: ncu --set full of the first conv (8 frames) after the taller-block restructure
cd /root/repo
mkdir -p gpurun_out
N=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"first_conv" -s 2 -c 1 -o gpurun_out/prof_first -f python scripts/mini_run.py > gpurun_out/ncu_first.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_first.log
