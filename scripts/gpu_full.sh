#!/bin/bash
# Full round check: GPU tests, smoke, bench (N=1), ncu launch list + full captures of the top kernels.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; grep smoke: gpurun_out/smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'], d['retrieval']['value'], d['retrieval']['roofline']['frac'], d['clocks'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1200 --csv --log-file gpurun_out/launches_tf32.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --skip-retrieval --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_wgrad_tc_kernel|bn_bwd_reg_kernel|bn_fwd_kernel" -s 1030 -c 30 -o gpurun_out/train_kernels -f \
  python bench.py --steps 2 --warmup 3 --no-graph --skip-retrieval --skip-cpu-baseline > gpurun_out/ncu_train.log 2>&1; echo "ncu train exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairwise_tc_kernel -s 1 -c 1 -o gpurun_out/pairwise_tc -f \
   python scripts/bench_pairwise.py 50000 100 1 tc > gpurun_out/ncu_pairwise.log 2>&1; echo "ncu pairwise exit $?"
