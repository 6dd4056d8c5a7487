#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "pairwise" > gpurun_out/pytest_pairwise.log 2>&1
echo "pairwise pytest exit $?"; tail -n 15 gpurun_out/pytest_pairwise.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -k "not pairwise" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['retrieval'])"; tail -n 5 gpurun_out/bench.err
