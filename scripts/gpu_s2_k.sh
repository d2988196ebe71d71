#!/bin/bash
# session 2, call K: cluster-multicast variant under sustained (power-capped) load
cd /root/repo
mkdir -p gpurun_out
for v in 0 1 0 1; do
  SQ_CLUSTER=$v timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s2k_cl$v.json 2> gpurun_out/bench_s2k_cl$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_s2k_cl$v.json'))
print('SQ_CLUSTER=$v value %.1f e2e %.1f clocks %s tensor frac %.3f' % (d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac']))"
done
