#!/bin/bash
timeout 100 python scripts/trace_bn_bwd.py 128 32 16; timeout 100 python scripts/trace_bn_bwd.py 128 8 64
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q --timeout=600 -k "bn or train or forward" 2>&1 | tail -n 5 | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_bn0.json 2> gpurun_out/bench_bn0.err
echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_bn0.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])
"; tail -n 3 gpurun_out/bench_bn0.err
