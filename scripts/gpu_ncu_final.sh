#!/bin/bash
# launch list + full capture of one forward pass at the bench batch (8 frames of 2048^2)
cd /root/repo
mkdir -p gpurun_out
export N=8
CMD="python scripts/mini_run.py"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit $?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_tc_kernel|conv_xc_kernel|first_conv' -s 44 -c 22 -o gpurun_out/prof_final -f $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu2.log
ls -la gpurun_out/ | tail -8
