"""Prints how the tcgen05 wgrad result relates to the fp32 kernel's (ratio per tap block)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_embeddings_b200 import _lib as L
L.load(); L.check(L.load().se_init())
N, H, C, Co = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (4, 32, 16, 16)
d = L.ConvDesc(N, H, H, C, Co, 3, 3, 1, 1, 1, H, H)
g = torch.Generator().manual_seed(0)
x = torch.randn(N, H, H, C, generator=g).cuda(); dy = torch.randn(N, H, H, Co, generator=g).cuda()
res = {}
for mode in (0, 1):
    dw = torch.zeros(3, 3, C, Co, device='cuda'); db = torch.zeros(Co, device='cuda')
    L.call('se_conv2d_wgrad', d, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), mode, L.stream_ptr())
    torch.cuda.synchronize()
    res[mode] = (dw.cpu().numpy(), db.cpu().numpy())
ref, got = res[0][0], res[1][0]
print('shape', (N, H, C, Co), 'ref |max|', np.abs(ref).max(), 'got |max|', np.abs(got).max(), 'nonzero frac', (got != 0).mean())
for tap in range(9):
    r, s = divmod(tap, 3)
    a, b = ref[r, s], got[r, s]
    print(' tap', tap, 'rel err %.3e' % (np.abs(a - b).max() / np.abs(a).max()), 'ratio median %.4f' % np.median(b[np.abs(a) > 0.1 * np.abs(a).max()] / a[np.abs(a) > 0.1 * np.abs(a).max()]))
print(' bias ref', res[0][1][:4], 'got', res[1][1][:4])
# does got match some other tap of ref? (tap / transpose confusion)
for tap in range(9):
    r, s = divmod(tap, 3)
    best = min(((np.abs(ref[r2, s2] - got[r, s]).max() / np.abs(ref).max(), (r2, s2)) for r2 in range(3) for s2 in range(3)))
    bestT = min(((np.abs(ref[r2, s2].T - got[r, s]).max() / np.abs(ref).max(), (r2, s2)) for r2 in range(3) for s2 in range(3))) if C == Co else (9, None)
    print(' got tap', (r, s), 'closest ref tap', best, 'closest transposed', bestT)
