#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q --timeout=600 -k "bn or train or forward or smoke" 2>&1 | tail -n 5 | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_val.json 2> gpurun_out/bench_val.err
echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_val.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['launches_per_step'], d['retrieval']['value'])
"; tail -n 3 gpurun_out/bench_val.err
