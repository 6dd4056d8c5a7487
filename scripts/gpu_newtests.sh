#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=400 -k "resnet50 or retrieval_api or pairwise_retrieval" 2>&1 | tail -n 15 | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json | cut -c1-3000; tail -n 5 gpurun_out/bench.err
