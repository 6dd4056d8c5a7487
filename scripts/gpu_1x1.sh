#!/bin/bash
# tensor-core coverage of the ResNet-50 shapes (1x1, 1x1/2, padded 3x3): unit parity, model tests, config-4 / headline bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "conv_fwd_dgrad_wgrad" 2>&1 | tail -n 25 | cut -c1-600
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=500 -k "resnet50 or wrn" 2>&1 | tail -n 8 | cut -c1-400
timeout 400 python bench.py --workload config4 --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_config4_tc.json 2> gpurun_out/bench_config4_tc.err
echo "config4 exit $?"; cut -c1-700 gpurun_out/bench_config4_tc.json; tail -n 3 gpurun_out/bench_config4_tc.err | cut -c1-300
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err
echo "headline exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_head.json')); print(d['value'], d['ms_per_step'])"; tail -n 2 gpurun_out/bench_head.err | cut -c1-300
timeout 300 python bench.py --workload config3 --steps 10 --warmup 3 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_config3_tc.json 2> gpurun_out/bench_config3_tc.err
echo "config3 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config3_tc.json')); print(d['value'], d['ms_per_step'])"
