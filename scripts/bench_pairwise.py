"""Times se_pairwise_dist (N=50000, D=100 by default) with CUDA events; prints Gpairs/s and HBM fraction."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_embeddings_b200 import _lib as L
from semantic_embeddings_b200.evaluate_retrieval import pairwise_distances

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mode = L.SE_MODE_TF32 if (len(sys.argv) <= 4 or sys.argv[4] == 'tc') else L.SE_MODE_F32
f = np.random.RandomState(0).randn(n, d).astype(np.float32)
f /= np.linalg.norm(f, axis=-1, keepdims=True)
fd = torch.from_numpy(f).cuda()
out = torch.empty((n, n), dtype=torch.float32, device='cuda')
for _ in range(2):
    pairwise_distances(None, False, 0, n, mode, out=out, feat_dev=fd)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, b in ev:
    a.record(); pairwise_distances(None, False, 0, n, mode, out=out, feat_dev=fd); b.record()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
best, mean = min(ms), float(np.mean(ms))
peak = 6567.4
byt = 4.0 * n * n + 4.0 * n * d
print(json.dumps({'N': n, 'D': d, 'ms_mean': mean, 'ms_best': best, 'gpairs_s': n * n / mean / 1e6,
                  'hbm_frac_mean': byt / mean / 1e6 / peak, 'hbm_frac_best': byt / best / 1e6 / peak}))
# spot check against fp64 on a few rows
ref = (f[:4].astype(np.float64) ** 2).sum(1)[:, None] + (f.astype(np.float64) ** 2).sum(1)[None, :] - 2 * f[:4].astype(np.float64) @ f.astype(np.float64).T
print('max abs err rows 0-3:', float(np.abs(out[:4].cpu().numpy() - ref).max()))
