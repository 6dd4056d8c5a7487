#!/bin/bash
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "softmax_corr or top_k_rank or sgd_schedule or xent or argsort or hier_metrics or augment or n50000 or embed_head" 2>&1 | tail -n 40 | cut -c1-400
echo "== models"
timeout 1200 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout=600 -k "batch128 or batch64 or resnet50_224 or cli_train" 2>&1 | tail -n 40 | cut -c1-500
cat gpurun_out/parity_models.jsonl | cut -c1-400
grep -h "n50000\|augment" gpurun_out/parity_ops.jsonl | cut -c1-300
