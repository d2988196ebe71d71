#!/bin/bash
# session 2, call G: sustained (power-capped) throughput of the conv variants; HBM traffic-mix probe
cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/hbm_probe.py > gpurun_out/hbm_probe.log 2>&1; cat gpurun_out/hbm_probe.log
for v in 1 2 0; do
  SQ_XC=$v timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s2g_xc$v.json 2> gpurun_out/bench_s2g_xc$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_s2g_xc$v.json'))
print('SQ_XC=$v value %.1f e2e %.1f clocks %s tensor frac %.3f' % (d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac']))"
done
