#!/bin/bash
# session 2, call I: frames per step (8 / 16 / 32) at equal frame counts, sustained regime
cd /root/repo
mkdir -p gpurun_out
for b in 8 16 32; do
  steps=$((480 / b))
  timeout 600 python bench.py --batch $b --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2i_b$b.json 2> gpurun_out/bench_s2i_b$b.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_s2i_b$b.json'))
print('batch $b steps $steps value %.1f e2e %.1f clocks %s tensor frac %.3f' % (d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac']))" || tail -3 gpurun_out/bench_s2i_b$b.err
done
