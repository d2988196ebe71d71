#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
SQ_DEBUG=1 timeout 300 python scripts/profile_layers.py > gpurun_out/layers_s2n_4.log 2>&1; grep "sequitr_b200:\|two-launch\|down0\|total" gpurun_out/layers_s2n_4.log
SQ_FUSE_GRID1=1 timeout 300 python scripts/profile_layers.py > gpurun_out/layers_s2n_g1.log 2>&1; grep "down0\|total" gpurun_out/layers_s2n_g1.log
