#!/bin/bash
# 2-GPU sanity at HEAD: data-parallel parity tests + the default torchrun bench line (what the driver's scaling run launches)
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=500 2>&1 | tail -n 12 | cut -c1-400
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench_n2_head.json 2> gpurun_out/r2_bench_n2_head.err
echo "bench n2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_head.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['gradient_exchange'], d['retrieval']['value'], d['n_gpus'])"; tail -n 3 gpurun_out/r2_bench_n2_head.err | cut -c1-300
