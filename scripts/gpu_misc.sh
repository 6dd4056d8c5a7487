#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=600 2>&1 | tail -n 6 | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_misc.json 2> gpurun_out/bench_misc.err
echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_misc.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step']); print(d['roofline'])
for b in d['breakdown']: print('  %-45s %3d  %.3f ms  %.1f us'%(b['kernel'],b['launches'],b['ms_per_step'],1000*b['ms_per_step']/b['launches']))
"; tail -n 3 gpurun_out/bench_misc.err
