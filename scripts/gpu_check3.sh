#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "conv or pairwise" > gpurun_out/pytest_conv.log 2>&1
echo "conv/pairwise pytest exit $?"; tail -n 30 gpurun_out/pytest_conv.log | cut -c1-300
timeout 300 python scripts/bench_pairwise.py 50000 100 10 tc | tee gpurun_out/pairwise_bench.txt
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=600 > gpurun_out/pytest_models.log 2>&1
echo "models pytest exit $?"; tail -n 12 gpurun_out/pytest_models.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step']); print(d['roofline']); [print(b) for b in d['breakdown'][:8]]; print(d['retrieval']['value'], d['retrieval']['roofline']['frac'])"; tail -n 5 gpurun_out/bench.err
