#!/bin/bash
# session 2, final check: full gpu suite, smoke, default bench, side kernels
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -6 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json'))
print('value %.1f e2e %.1f e2e_u16 %.1f clocks %s roofline %.3f hbm %.3f floor %.3f cpu %.2f launches %d' % (d['value'], d['e2e']['value'], d['e2e_raw_u16']['value'], d['clocks'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['roofline_hbm']['layer_floor']['frac'], d['cpu_baseline']['value'], d['gpu_launches']))"; tail -2 gpurun_out/bench_final.err
timeout 600 python scripts/bench_aux.py > gpurun_out/aux_final.log 2>&1; grep -v '^{' gpurun_out/aux_final.log | tail -10
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_final.json 2> gpurun_out/ref_final.err; echo "reference exit $?"; cut -c1-300 gpurun_out/ref_final.json
