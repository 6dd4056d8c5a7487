"""clock64 timeline of CTA 0 of the fused conv + BatchNorm kernel (SE_CT_TRACE_PTR debug knob)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trace = torch.zeros(3 * 512, dtype=torch.int64, device='cuda')
os.environ['SE_CT_TRACE_PTR'] = str(trace.data_ptr())
from semantic_embeddings_b200 import _lib as L
L.load(); L.check(L.load().se_init())
N, H, C, Co = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (128, 32, 16, 16)
d = L.ConvDesc(N, H, H, C, Co, 3, 3, 1, 1, 1, H, H)
x = torch.randn(N, H, H, C, device='cuda'); w = torch.randn(3, 3, C, Co, device='cuda') * 0.1
wt = torch.empty_like(w)
tab = (ctypes.c_int64 * 4)(0, 9, C, Co)
L.call('se_transpose_filters', w.data_ptr(), wt.data_ptr(), tab, 1, L.stream_ptr())
y = torch.empty(N, H, H, Co, device='cuda'); z = torch.empty_like(y)
g = torch.ones(Co, device='cuda'); b = torch.zeros(Co, device='cuda'); mm = torch.zeros(Co, device='cuda'); mv = torch.ones(Co, device='cuda')
sm = torch.zeros(Co, device='cuda'); si = torch.zeros(Co, device='cuda')
for _ in range(3):
    stats = torch.zeros(2 * Co, dtype=torch.float64, device='cuda'); cnt = torch.zeros(1, dtype=torch.int64, device='cuda')
    L.call('se_conv_bn_fwd', d, x.data_ptr(), w.data_ptr(), wt.data_ptr(), b.data_ptr(), y.data_ptr(), 0, stats.data_ptr(),
           g.data_ptr(), b.data_ptr(), 1e-3, 0.99, mm.data_ptr(), mv.data_ptr(), sm.data_ptr(), si.data_ptr(), None, 1, z.data_ptr(),
           cnt.data_ptr(), 1, L.stream_ptr())
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(3, 256, 2)
t0 = min(int(t[r, 0, 1]) for r in range(3) if t[r, 0, 1] > 0)
names = {0: {0: 'start', 1: 'got-empty', 2: 'tma-issued'}, 1: {0: 'start', 1: 'got-tmem-empty', 2: 'got-full', 3: 'committed'},
         2: {0: 'start', 1: 'got-tmem-full', 2: 'tmem-ld-done', 3: 'tile-done', 4: 'pass1-all-warps', 5: 'stats-atomics-done',
             6: 'grid-barrier-passed', 7: 'coef-ready', 8: 'pass2-done'}}
for r, role in enumerate(('producer', 'mma', 'epilogue')):
    ev = [(int(e), int(c) - t0) for e, c in t[r][:200] if c > 0]
    print(role, len(ev), 'events')
    print('   ', ' '.join('%s@%d' % (names[r].get(e, str(e)), c) for e, c in ev[:60]))
print('first block of the first tile (cycles from entry): combine %d, stores %d, stats %d' % (
    int(t[2, 200, 1] - t[2, 200, 0]), int(t[2, 201, 0] - t[2, 200, 1]), int(t[2, 201, 1] - t[2, 201, 0])))
