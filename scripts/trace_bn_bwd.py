"""clock64 phase stamps of CTA 0 of bn_bwd_reg_kernel (SE_BN_TRACE_PTR debug knob)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trace = torch.zeros(8, dtype=torch.int64, device='cuda')
os.environ['SE_BN_TRACE_PTR'] = str(trace.data_ptr())
from semantic_embeddings_b200 import _lib as L
L.load(); L.check(L.load().se_init())
N, H, C = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (128, 32, 16)
rows = N * H * H
x = torch.randn(rows, C, device='cuda'); y = torch.relu(torch.randn(rows, C, device='cuda')); d = torch.randn(rows, C, device='cuda')
g = torch.ones(C, device='cuda'); sm = torch.zeros(C, device='cuda'); si = torch.ones(C, device='cuda')
dx = torch.empty_like(x); dres = torch.empty_like(x); dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda')
big = torch.empty(64 << 20, device='cuda')
for it in range(3):
    big.zero_()                                   # evict x / y / d from L2: the in-step situation for x and y
    d2 = d.clone()                                 # ... while dout is fresh in L2
    scratch = torch.zeros(2 * C + 1, dtype=torch.float64, device='cuda')
    L.call('se_bn_bwd', x.data_ptr(), y.data_ptr(), d2.data_ptr(), rows, C, g.data_ptr(), sm.data_ptr(), si.data_ptr(), 1, 0,
           dx.data_ptr(), 0.0, dres.data_ptr(), 0.0, dg.data_ptr(), db.data_ptr(), scratch.data_ptr(), L.stream_ptr())
    torch.cuda.synchronize()
t = trace.cpu().numpy()
names = ['start', 'x,y loads issued', 'after pdl_wait', 'dout + local sums', 'global atomics + fence', 'grid barrier passed', 'coefficients', 'stores issued']
print('C=%d rows=%d' % (C, rows), ' '.join('%s@%d' % (n, int(v - t[0])) for n, v in zip(names, t)))
