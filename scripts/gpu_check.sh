#!/bin/bash
# One GPU-box session: parity tests, smoke, a short bench, and the ncu launch list of the bench.
# Usage (from the repo root, via gpurun): bash scripts/gpu_check.sh [tests|bench|ncu|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == all || $what == tests ]]; then
  rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
  timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -n 40 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -n 5 gpurun_out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -c 6000 gpurun_out/bench.json; tail -n 20 gpurun_out/bench.err
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --skip-retrieval --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_bench.log
fi
