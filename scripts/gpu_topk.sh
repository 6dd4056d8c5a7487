#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=400 -x -k "fused or n50000 or pairwise" 2>&1 | tail -n 25 | cut -c1-400
timeout 500 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r.json')); r=d['retrieval']; print(r['value'], r['ms'], r['roofline']['frac']); print(r.get('fused_topk')); print(r.get('ranking_top251')); print(r.get('ranking_full')); print(r.get('metrics_full'))"; tail -n 3 gpurun_out/bench_r.err | cut -c1-300
