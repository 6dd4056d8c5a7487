#!/bin/bash
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_ops.jsonl
timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "conv_fwd_dgrad_wgrad and tf32x3" 2>&1 | tail -n 4 | cut -c1-600
run() {
  timeout 300 python bench.py --mode tf32x3 --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  echo "bench [$1] exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print(d['value'], d['ms_per_step']); print(' | '.join('%s %.3f' % (b['kernel'].replace('conv_','').replace(' 3x3 s1',''), b['ms_per_step']) for b in d['breakdown'][:12]))"; tail -n 2 gpurun_out/bench_x.err
}
run default
SE_CT_NO_COOP_GENERIC=1 run no_coop_generic
SE_WG_NO_PT32=1 run no_pt32
SE_WG_NO_PT32=1 SE_CT_NO_COOP_GENERIC=1 run neither
cp gpurun_out/bench_x.json gpurun_out/bench_tf32x3.json
