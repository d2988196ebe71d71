#!/bin/bash
# session 2, call F: is the e2e gap a clock effect?  long device-resident run vs the default
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s2f_long.json 2> gpurun_out/bench_s2f_long.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_s2f_long.json'))
print('value',d['value'],'e2e',d['e2e'],'clocks',d['clocks'],'floor',d['roofline_hbm'].get('layer_floor'))"
