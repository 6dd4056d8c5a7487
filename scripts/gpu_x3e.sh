#!/bin/bash
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_ops.jsonl
timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "conv_fwd_dgrad_wgrad or conv_bn or conv_epilogue" 2>&1 | tail -n 6 | cut -c1-600
echo "== layer times (us, warm)"
timeout 200 python scripts/bench_conv.py 2>&1 | grep shape
SE_CT_NO_SINGLE=1 timeout 200 python scripts/bench_conv.py 2>&1 | grep shape | head -2
for m in tf32x3 tf32; do
  timeout 300 python bench.py --mode $m --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  echo "bench $m exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$m.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step'], d['e2e']['value']); print(' | '.join('%s %.3f' % (b['kernel'].replace('conv_','').replace(' 3x3 s1',''), b['ms_per_step']) for b in d['breakdown'][:12]))"; tail -n 2 gpurun_out/bench_$m.err
done
