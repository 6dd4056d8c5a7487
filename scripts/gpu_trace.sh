#!/bin/bash
timeout 120 python scripts/trace_conv.py 128 32 16 16
timeout 120 python scripts/trace_conv.py 128 8 64 64
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=200 -k "conv" 2>&1 | tail -n 8 | cut -c1-300
env SE_CT_DEBUG=0 timeout 300 python scripts/bench_conv.py 2>&1 | grep shape | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['shape'], d['us'])
"
