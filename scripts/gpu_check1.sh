#!/bin/bash
# first GPU pass: tcgen05/TMA convention probes, then the gpu parity tests
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/probe.log 2>&1
for c in gemm gemm256 conv16 conv64 tma; do
  timeout 60 sequitr_b200/_build/tc_probe $c >> gpurun_out/probe.log 2>&1
  echo "probe $c exit $?" >> gpurun_out/probe.log
done
grep -E "VERDICT|exit|error|timed" gpurun_out/probe.log
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"
tail -40 gpurun_out/pytest_gpu.log
