#!/bin/bash
mkdir -p gpurun_out
for G in 1 2 4 8; do
  if [ "$G" -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$G.json 2> gpurun_out/scale_$G.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29520+G)) \
       bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/scale_$G.json 2> gpurun_out/scale_$G.err
  fi
  echo "G=$G exit $?"; python -c "import json;d=json.loads(open('gpurun_out/scale_$G.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['e2e']['value'],d['e2e_raw_u16']['value'],d['ms_per_step'],d['clocks'])"
done
