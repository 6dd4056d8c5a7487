#!/bin/bash
# 2-GPU checks: data-parallel parity (native NCCL-in-graph and torch paths), then bench at N=2 for both, weak and strong
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=500 2>&1 | tail -n 25 | cut -c1-600
for c in native torch; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --skip-retrieval --comm $c > gpurun_out/bench_n2_$c.json 2> gpurun_out/bench_n2_$c.err
  echo "bench n2 $c exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_$c.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['gradient_exchange'])"; tail -n 3 gpurun_out/bench_n2_$c.err | cut -c1-300
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --skip-retrieval --scaling strong > gpurun_out/bench_n2_strong.json 2> gpurun_out/bench_n2_strong.err
echo "strong exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_strong.json')); print(d['value'], d['ms_per_step'], d['scaling'], d['config']['per_gpu_batch'])"
timeout 300 python bench.py --steps 20 --warmup 5 --skip-retrieval --skip-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('n1', d['value'], d['ms_per_step'])"
