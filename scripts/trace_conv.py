"""Dumps the per-role clock64 timeline of CTA 0 of conv_tc_kernel (debug build knob SE_CT_TRACE_PTR)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trace = torch.zeros(5 * 512, dtype=torch.int64, device='cuda')
os.environ['SE_CT_TRACE_PTR'] = str(trace.data_ptr())
from semantic_embeddings_b200 import _lib as L
L.load(); L.check(L.load().se_init())
N, H, C, Co = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (128, 32, 16, 16)
d = L.ConvDesc(N, H, H, C, Co, 3, 3, 1, 1, 1, H, H)
x = torch.randn(N, H, H, C, device='cuda'); w = torch.randn(3, 3, C, Co, device='cuda') * 0.1
y = torch.empty(N, H, H, Co, device='cuda'); dx = torch.empty_like(x)
mode = int(os.environ.get('TRACE_MODE', '2'))
wt, wl, wtl = torch.empty_like(w), torch.empty_like(w), torch.empty_like(w)
tab = (ctypes.c_int64 * 4)(0, 9, C, Co)
L.call('se_split_filters', w.data_ptr(), wt.data_ptr(), wl.data_ptr(), wtl.data_ptr(), tab, 1, L.stream_ptr())
aux = L.ConvAux(wt.data_ptr(), wtl.data_ptr(), wl.data_ptr())
for _ in range(3):
    L.call('se_conv2d_dgrad_aux', d, y.data_ptr(), w.data_ptr(), aux, dx.data_ptr(), 0.0, mode, L.stream_ptr())
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(5, 256, 2)
t0 = min(int(t[r, 0, 1]) for r in range(3) if t[r, 0, 1] > 0)
print('mode', mode, 'shape', N, H, C, Co)
names = {0: {0: 'start', 1: 'got-empty', 2: 'tma-issued'},
         1: {0: 'start', 1: 'got-tmem-empty', 2: 'got-full', 3: 'committed', 4: 'got-lo', 5: 'committed2'},
         2: {0: 'start', 1: 'got-tmem-full', 2: 'tmem-ld-done', 3: 'tile-done'}, 3: {1: 'got-hi-done', 2: 'split-done'}, 4: {4: 'got-lo', 5: 'committed2'}}
for r, role in enumerate(('producer', 'mma', 'epilogue', 'splitter', 'mma2')):
    ev = [(int(e), int(c) - t0) for e, c in t[r][:200] if c > 0]
    print(role, len(ev), 'events')
    print('   ', ' '.join('%s@%d' % (names[r].get(e, str(e)), c) for e, c in ev[:60]))
