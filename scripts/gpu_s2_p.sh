#!/bin/bash
# session 2, call P: 8-rank run of the default bench (torchrun, one rank per GPU, no collective on the data path)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/scale_s2_8.json 2> gpurun_out/scale_s2_8.err
echo "exit $?"; python -c "import json;d=json.loads(open('gpurun_out/scale_s2_8.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['e2e']['value'],d['e2e_raw_u16']['value'],d['ms_per_step'],d['clocks'])"; tail -2 gpurun_out/scale_s2_8.err
