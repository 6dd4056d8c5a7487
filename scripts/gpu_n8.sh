#!/bin/bash
# 8-GPU bench lines: weak scaling with the library's NCCL-in-graph exchange and with torch.distributed, strong scaling
cd "$(dirname "$0")/.."
for c in native torch; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 30 --warmup 5 --comm $c > gpurun_out/bench_n8_$c.json 2> gpurun_out/bench_n8_$c.err
  echo "bench n8 $c exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n8_$c.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['gradient_exchange'][:40], d['retrieval']['value'] if d.get('retrieval') else None)"; grep -i "error\|Traceback" gpurun_out/bench_n8_$c.err | head -5
done
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 30 --warmup 5 --skip-retrieval --scaling strong > gpurun_out/bench_n8_strong.json 2> gpurun_out/bench_n8_strong.err
echo "strong exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n8_strong.json')); print(d['value'], d['ms_per_step'], d['scaling'], d['config']['per_gpu_batch'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --steps 10 --warmup 3 --skip-retrieval --workload config3 > gpurun_out/bench_n8_config3.json 2> gpurun_out/bench_n8_config3.err
echo "config3 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n8_config3.json')); print(d['value'], d['ms_per_step'], d['config']['workload'][:50], d['conv_flop_roofline'])"
