#!/bin/bash
# full gpu suite + smoke + bench + other configs + aux kernels
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.log 2>&1; grep -v '^\[' gpurun_out/configs.log
timeout 600 python scripts/bench_aux.py > gpurun_out/aux.log 2>&1; grep -v '^{' gpurun_out/aux.log
timeout 600 python scripts/bench_config4.py > gpurun_out/config4.log 2>&1
echo "c4 exit $?"; grep -v '^{' gpurun_out/config4.log
timeout 300 python scripts/profile_layers.py > gpurun_out/layers.log 2>&1; cat gpurun_out/layers.log
