#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout=900 2>&1 | tail -n 6 | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_chk.json')); print(d['value'], d['ms_per_step'], d['dtype'], d['launches_per_step'], d['e2e']['value'], d['clocks'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
