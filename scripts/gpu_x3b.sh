#!/bin/bash
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_ops.jsonl
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -x -k "conv_fwd_dgrad_wgrad and tf32x3" 2>&1 | tail -n 5 | cut -c1-400
echo "== layer times (us, warm)"
timeout 200 python scripts/bench_conv.py 2>&1 | grep shape
SE_WG_NO_PACK=1 timeout 200 python scripts/bench_conv.py 2>&1 | grep shape | head -1
echo "== traces"
for sh in "128 32 16 16" "128 16 32 32" "128 8 64 64"; do
  TRACE_MODE=2 timeout 100 python scripts/trace_conv.py $sh 2>&1 | cut -c1-1500
  TRACE_MODE=1 timeout 100 python scripts/trace_conv.py $sh 2>&1 | cut -c1-900
done
for m in tf32x3; do
  timeout 300 python bench.py --mode $m --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  echo "bench $m exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$m.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step'], d['e2e']['value']); [print(b) for b in d['breakdown'][:14]]"; tail -n 3 gpurun_out/bench_$m.err
done
