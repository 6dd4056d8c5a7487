#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q --timeout=600 -k "bn or train or forward" 2>&1 | tail -n 5 | cut -c1-300
for e in 0 1; do
  if [ $e = 1 ]; then export SE_BN_NO_REG=1; else unset SE_BN_NO_REG; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_bn$e.json 2> gpurun_out/bench_bn$e.err
  echo "NO_REG=$e exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_bn$e.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])
for b in d['breakdown']:
    if 'bn' in b['kernel']: print('  %-45s %3d  %.3f ms  %.1f us'%(b['kernel'],b['launches'],b['ms_per_step'],1000*b['ms_per_step']/b['launches']))
"; tail -n 3 gpurun_out/bench_bn$e.err
done
