"""Component times of the fused distance + top-k pipeline (ncu-free: CUDA events around repeated calls)."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_embeddings_b200 import _lib as L
from semantic_embeddings_b200.evaluate_retrieval import pairwise_distances, pairwise_topk, row_topk
n, d, k = 50000, 100, 251
f = np.random.RandomState(0).randn(n, d).astype(np.float32); f /= np.linalg.norm(f, axis=-1, keepdims=True)
fd = torch.from_numpy(f).cuda()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print('fused total', t(lambda: pairwise_topk(k=k, feat_dev=fd)))
samp = torch.empty((n, 4096), device='cuda')
print('row_topk on [N,4096] k=78', t(lambda: row_topk(samp.normal_(), 78)), '(incl. normal_)', t(lambda: samp.normal_()))
out = torch.empty((8192, n), device='cuda')
print('dist 8192 rows', t(lambda: pairwise_distances(None, False, 0, 8192, 2, out=out, feat_dev=fd)))
