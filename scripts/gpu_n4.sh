#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "n4 exit $?"; wc -l gpurun_out/bench_n4.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['retrieval']['value'], d['retrieval'].get('ranking_top251'))"
