#!/bin/bash
cd "$(dirname "$0")/.."
timeout 120 python bench.py --batch 16 --steps 10 --warmup 3 --skip-retrieval --skip-cpu-baseline > gpurun_out/b16_n1.json 2> gpurun_out/b16_n1.err
echo "n1 b16 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/b16_n1.json')); print(d['value'], d['ms_per_step'])"
for c in torch native; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --skip-retrieval --batch 32 --scaling strong --comm $c > gpurun_out/b16_n2_$c.json 2> gpurun_out/b16_n2_$c.err
  echo "n2 b16 $c exit $?"; python -c "
import json; d=json.load(open('gpurun_out/b16_n2_$c.json')); print(d['value'], d['ms_per_step'], d['config']['per_gpu_batch'])"
done
