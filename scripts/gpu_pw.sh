#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=200 -k "pairwise" 2>&1 | tail -n 6 | cut -c1-400
timeout 200 python scripts/bench_pairwise.py 50000 100 10 tc
timeout 200 python scripts/bench_pairwise.py 10000 100 10 tc
