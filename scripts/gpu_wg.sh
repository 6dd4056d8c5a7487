#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "conv" 2>&1 | tail -n 12 | cut -c1-300
for dbg in 0 1 4 7; do
SE_WG_DEBUG=$dbg timeout 300 python scripts/bench_conv.py 2>&1 | grep shape | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('dbg=$dbg', d['shape'], d['us'])
"
done
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); [print(b) for b in d['breakdown'][:14]]"; tail -n 5 gpurun_out/bench.err
