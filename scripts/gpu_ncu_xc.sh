#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
export N=2
CMD="python scripts/profile_layers.py"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_xc_kernel' -s 12 -c 4 \
    -o gpurun_out/prof_xc -f $CMD > gpurun_out/ncu_xc.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_xc.log
