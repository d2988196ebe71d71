#!/bin/bash
# session 2, call O: full gpu suite with the opt-in fusion test, layer times with the default path
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/profile_layers.py > gpurun_out/layers_s2o.log 2>&1; grep "down0\|total" gpurun_out/layers_s2o.log
SQ_FUSE_FIRST=1 timeout 300 python scripts/profile_layers.py > gpurun_out/layers_s2o_fused.log 2>&1; grep "down0\|total" gpurun_out/layers_s2o_fused.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
