#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q --timeout=600 -k "bn or train or forward" 2>&1 | tail -n 5 | cut -c1-300
for e in 1 0; do
  if [ $e = 1 ]; then export SE_NO_CONV_BN_FUSION=1; else unset SE_NO_CONV_BN_FUSION; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_fuse$e.json 2> gpurun_out/bench_fuse$e.err
  echo "NO_FUSE=$e exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_fuse$e.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])
for b in d['breakdown']:
    if 'bn' in b['kernel']: print('  %-45s %3d  %.3f ms  %.1f us'%(b['kernel'],b['launches'],b['ms_per_step'],1000*b['ms_per_step']/b['launches']))
"; tail -n 3 gpurun_out/bench_fuse$e.err
done
