#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_label.py tests/test_gpu_unet_bf16.py tests/test_gpu_unet_fp32.py -m gpu -q --no-header -rf --timeout 300 -x > gpurun_out/pytest_sel.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_sel.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 16 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  CMD="python scripts/profile_layers.py"
  export N=2
  timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:"conv_tc_kernel<16, 4|conv_tc_kernel<32, 4|first_conv" -s 14 -c 7 \
      -o gpurun_out/prof_l0 -f $CMD > gpurun_out/ncu3.log 2>&1
  echo "ncu exit $?"; tail -3 gpurun_out/ncu3.log
fi
