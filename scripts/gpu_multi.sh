#!/bin/bash
# 2-GPU data-parallel run of bench.py exactly as the driver launches it
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 exit $?"; tail -c 1500 gpurun_out/bench_n2.json; tail -n 8 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref n2 exit $?"; tail -c 600 gpurun_out/bench_ref_n2.json
timeout 600 python scripts/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check exit $?"; tail -n 5 gpurun_out/dp_check.log
