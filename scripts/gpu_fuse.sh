#!/bin/bash
timeout 120 python scripts/trace_conv_bn.py 128 32 16 16 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 2>&1 | tail -n 6 | cut -c1-300
timeout 300 python scripts/bench_conv.py 2>&1 | grep shape | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['shape'], d['us'])
"
for e in 0 1; do
  if [ $e = 1 ]; then export SE_NO_CONV_BN_FUSION=1; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_fuse$e.json 2> gpurun_out/bench_fuse$e.err
  echo "NO_FUSE=$e exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_fuse$e.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"; tail -n 3 gpurun_out/bench_fuse$e.err
done
