#!/bin/bash
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_ops.jsonl
timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "conv_fwd_dgrad_wgrad or truncated" 2>&1 | tail -n 8 | cut -c1-600
echo "== layer times (us, warm)"
timeout 200 python scripts/bench_conv.py 2>&1 | grep shape
for m in tf32x3 tf32; do
  timeout 300 python bench.py --mode $m --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  echo "bench $m exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$m.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step'], d['e2e']['value']); [print(b['kernel'], round(b['ms_per_step'],3)) for b in d['breakdown'][:14]]"; tail -n 3 gpurun_out/bench_$m.err
done
