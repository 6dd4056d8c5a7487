#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=200 -k "pairwise" 2>&1 | tail -3
timeout 300 python scripts/bench_pairwise.py 50000 100 10 tc | tee gpurun_out/pairwise_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairwise_tc_kernel -s 1 -c 1 -o gpurun_out/pairwise_tc -f \
   python scripts/bench_pairwise.py 50000 100 1 tc > gpurun_out/ncu_pairwise.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_pairwise.log
