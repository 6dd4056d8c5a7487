#!/bin/bash
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
timeout 2400 python -m pytest tests/ -x -q -m gpu --timeout=900 --durations=12 2>&1 | tail -n 40 | cut -c1-300
for w in config3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step'], d['e2e']['value'], d['conv_flop_roofline']); [print(b['kernel'], round(b['ms_per_step'],3), b['launches']) for b in d['breakdown'][:10]]"; tail -n 3 gpurun_out/bench_$w.err | cut -c1-300
done
