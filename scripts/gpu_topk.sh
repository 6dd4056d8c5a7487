#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout=300 -k "topk or pairwise" 2>&1 | tail -n 8 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_topk.json 2> gpurun_out/bench_topk.err
echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_topk.json')); print(d['value'], d['ms_per_step']); print(d['retrieval']['value'], d['retrieval'].get('ranking_top251'))
"; tail -n 3 gpurun_out/bench_topk.err
