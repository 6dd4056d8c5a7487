#!/bin/bash
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "n8 exit $?"; wc -l gpurun_out/bench_n8.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['retrieval']['value'])"; tail -n 4 gpurun_out/bench_n8.err | cut -c1-300
