#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=400 -x -k "fused" 2>&1 | tail -n 8 | cut -c1-400
timeout 300 python scripts/time_topk.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/topk_launches.csv python -c "
import sys; sys.path.insert(0,'.')
import numpy as np, torch
from semantic_embeddings_b200.evaluate_retrieval import pairwise_topk
f=np.random.RandomState(0).randn(50000,100).astype(np.float32); f/=np.linalg.norm(f,axis=-1,keepdims=True)
fd=torch.from_numpy(f).cuda()
pairwise_topk(k=251, feat_dev=fd); torch.cuda.synchronize()
" > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/topk_launches.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]; hdr=rows[h]; ix={n:i for i,n in enumerate(hdr)}
for r in rows[h+1:]:
    if len(r)>=len(hdr) and r[ix['Metric Name']]=='gpu__time_duration.sum':
        print(r[ix['Kernel Name']][:60], r[ix['Metric Value']], r[ix['Metric Unit']])
PY
