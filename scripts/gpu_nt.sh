#!/bin/bash
for nt in 4 2 1; do
SE_CT_NT=$nt timeout 300 python scripts/bench_conv.py 2>&1 | grep shape | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); u=d['us']; print('nt=$nt', d['shape'], u['fwd_tc'], u['dgrad_tc'])
"
done
for nt in 2 1; do
SE_CT_NT=$nt SE_NO_CONV_BN_FUSION=1 timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-retrieval > gpurun_out/bench_nt$nt.json 2> gpurun_out/bench_nt$nt.err
python -c "
import json; d=json.load(open('gpurun_out/bench_nt$nt.json')); print('nt=$nt', d['value'], d['ms_per_step'])"
done
