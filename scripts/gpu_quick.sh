#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=200 -k "conv or bn" 2>&1 | tail -n 6 | cut -c1-300
timeout 300 python scripts/bench_conv.py 2>&1 | grep shape | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['shape'], d['env'], d['us'])
"
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step']); [print(b) for b in d['breakdown'][:12]]"; tail -n 5 gpurun_out/bench.err
