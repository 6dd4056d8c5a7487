#!/bin/bash
# first check of the error-compensated (tf32x3) kernels: op parity, hardware-truncation probe, step parity, bench lines
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -x -k "truncated or (conv_fwd_dgrad_wgrad and tf32x3)" 2>&1 | tail -n 15 | cut -c1-400
echo "== models"
timeout 500 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=400 -k "simple-x3 or resnet-110-fc-x3" 2>&1 | tail -n 15 | cut -c1-600
for m in tf32x3 tf32; do
  timeout 300 python bench.py --mode $m --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  echo "bench $m exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$m.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step'], d['e2e']['value']); [print(b) for b in d['breakdown'][:14]]"; tail -n 3 gpurun_out/bench_$m.err
done
cat gpurun_out/parity_ops.jsonl | tail -n 60 | cut -c1-300
