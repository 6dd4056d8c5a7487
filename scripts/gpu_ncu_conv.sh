#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1400 --csv --log-file gpurun_out/launches_tc.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --skip-retrieval --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 300 -c 2 -o gpurun_out/conv_tc -f \
  python bench.py --steps 2 --warmup 3 --no-graph --skip-retrieval --skip-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_conv.log
