#!/bin/bash
# full gpu test suite, bench line, ncu launch list + one full capture of the conv kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 1 --batch 2 --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
  echo "ncu launches exit $?"
  timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 21 \
      -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu2.log 2>&1
  echo "ncu full exit $?"
  ls -la gpurun_out/
fi
