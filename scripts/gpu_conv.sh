#!/bin/bash
timeout 100 python scripts/trace_conv.py 128 8 64 64 | cut -c1-700
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q --timeout=600 -k "conv or train or forward" 2>&1 | tail -n 5 | cut -c1-300
timeout 300 python scripts/bench_conv.py 2>&1 | grep shape | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); u=d['us']; print(d['shape'], u['fwd_tc'], u['dgrad_tc'], u['wgrad_tc'])
"
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_conv.json 2> gpurun_out/bench_conv.err
echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_conv.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])
"; tail -n 3 gpurun_out/bench_conv.err
