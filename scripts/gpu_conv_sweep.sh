#!/bin/bash
set -u
mkdir -p gpurun_out; : > gpurun_out/conv_sweep.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "conv or bn" 2>&1 | tail -n 12 | cut -c1-250
for env in "" "SE_CT_DEBUG=1" "SE_CT_DEBUG=2" "SE_CT_DEBUG=4" "SE_CT_STAGES=2"; do
  env $env timeout 300 python scripts/bench_conv.py 2>&1 | grep shape >> gpurun_out/conv_sweep.txt
done
python -c "
import sys, json
for l in open('gpurun_out/conv_sweep.txt'):
    d = json.loads(l); print(d['shape'], d['env'], d['us'])
"
