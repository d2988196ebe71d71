#!/bin/bash
# session 2, call H: two-rank sanity run of both bench arms (torchrun, one rank per GPU)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/scale_s2_2.json 2> gpurun_out/scale_s2_2.err
echo "ours exit $?"; python -c "import json;d=json.loads(open('gpurun_out/scale_s2_2.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['e2e']['value'],d['e2e_raw_u16']['value'],d['ms_per_step'],d['clocks'])"; tail -2 gpurun_out/scale_s2_2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
   bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/ref_s2_2.json 2> gpurun_out/ref_s2_2.err
echo "reference exit $?"; cat gpurun_out/ref_s2_2.json | cut -c1-600
