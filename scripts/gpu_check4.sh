#!/bin/bash
set -u
mkdir -p gpurun_out; : > gpurun_out/conv_sweep.txt
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -n 12 | cut -c1-250
for env in ""; do
  env $env timeout 300 python scripts/bench_conv.py 2>&1 | grep shape >> gpurun_out/conv_sweep.txt
done
python -c "
import sys, json
for l in open('gpurun_out/conv_sweep.txt'):
    d = json.loads(l); print(d['shape'], d['env'], d['us'])
"
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step']); print(d['roofline']); [print(b) for b in d['breakdown'][:10]]; print(d['retrieval']['value'], d['retrieval']['roofline']['frac'])"; tail -n 5 gpurun_out/bench.err
