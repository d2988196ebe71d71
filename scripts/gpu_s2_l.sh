#!/bin/bash
# session 2, call L: CCL merge restricted to one segment + 4-wide scan; cluster variant under sustained load
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_label.py tests/test_gpu_unet_bf16.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/pytest_s2l.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_s2l.log
timeout 600 python scripts/bench_aux.py > gpurun_out/aux_s2l.log 2>&1; grep "label_centroids" gpurun_out/aux_s2l.log | grep -v '^{'
for v in 0 1; do
  SQ_CLUSTER=$v timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s2k_cl$v.json 2> gpurun_out/bench_s2k_cl$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_s2k_cl$v.json'))
print('SQ_CLUSTER=$v value %.1f e2e %.1f clocks %s tensor frac %.3f' % (d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac']))"
done
