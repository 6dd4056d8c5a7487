#!/bin/bash
for e in 0 1; do
  if [ $e = 1 ]; then export SE_NO_PDL=1; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_pdl$e.json 2> gpurun_out/bench_pdl$e.err
  echo "NO_PDL=$e exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$e.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -n 3 gpurun_out/bench_pdl$e.err
done
unset SE_NO_PDL
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-retrieval --no-graph > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_bench.log | cut -c1-300
