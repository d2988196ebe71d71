#!/bin/bash
# full gpu suite + bench (+ optional multi-GPU bench when GPUS>1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
G=${GPUS:-1}
if [ "$G" -gt 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_$G.json 2> gpurun_out/bench_$G.err
  echo "bench x$G exit $?"; cat gpurun_out/bench_$G.json; tail -3 gpurun_out/bench_$G.err
fi
