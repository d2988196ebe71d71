#!/bin/bash
# session 2, call C: ncu --set full of the side kernels (second pass of scripts/aux_run.py)
cd /root/repo
mkdir -p gpurun_out
CMD="python scripts/aux_run.py"
timeout 300 $CMD > gpurun_out/aux_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"norm_|bg_|augment|outliers|inst_|edt_|run_|scan_|root_|ccl_" -c 60 -o gpurun_out/prof_aux_s2 -f $CMD > gpurun_out/ncu_aux_s2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_aux_s2.log; tail -2 gpurun_out/aux_plain.log
