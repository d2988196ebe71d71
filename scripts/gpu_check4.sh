#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_label.py tests/test_gpu_weightmap.py tests/test_gpu_unet_fp32.py -m gpu -q --no-header -rf --timeout 300 > gpurun_out/pytest_sel.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_sel.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
