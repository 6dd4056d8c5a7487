#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 2>&1 | tail -n 6 | cut -c1-300
for e in 0 1; do
  if [ $e = 1 ]; then export SE_NO_SIDE_STREAM=1; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_side$e.json 2> gpurun_out/bench_side$e.err
  echo "NO_SIDE=$e exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_side$e.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -n 3 gpurun_out/bench_side$e.err
done
