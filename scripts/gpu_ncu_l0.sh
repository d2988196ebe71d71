#!/bin/bash
mkdir -p gpurun_out
export N=2
CMD="python scripts/profile_layers.py"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_tc_kernel<\(int\)16, \(int\)4|conv_tc_kernel<\(int\)32, \(int\)4|first_conv|conv_tc_kernel<\(int\)16, \(int\)2' -s 18 -c 9 \
    -o gpurun_out/prof_l0 -f $CMD > gpurun_out/ncu3.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu3.log
