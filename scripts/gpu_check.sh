#!/bin/bash
# full GPU test run + short bench lines (headline, config 4) + smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu --timeout=900 2>&1 | tail -n 25 | cut -c1-400
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_chk.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step'], d['e2e']['value'], d['clocks'])"
timeout 400 python bench.py --workload config4 --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_config4_chk.json 2> gpurun_out/bench_config4_chk.err
echo "config4 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config4_chk.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac']); [print(b) for b in d['breakdown'][:8]]"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
