#!/bin/bash
rm -f gpurun_out/parity_ops.jsonl gpurun_out/parity_models.jsonl
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; grep smoke: gpurun_out/smoke.log
