"""Times one conv layer (fwd / dgrad / wgrad) per arithmetic mode with CUDA events (50 launches replayed from a graph)."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_embeddings_b200 import _lib as L

MODES = {'x3': 2, 'tc': 1, 'f32': 0}


def bench(N, H, C, Co, k=3, reps=50, modes=('x3', 'tc')):
    L.load(); L.check(L.load().se_init())
    d = L.ConvDesc(N, H, H, C, Co, k, k, 1, k // 2, k // 2, H, H)
    x = torch.randn(N, H, H, C, device='cuda'); w = torch.randn(k, k, C, Co, device='cuda') * 0.1
    wt, wl, wtl = torch.empty_like(w), torch.empty_like(w), torch.empty_like(w)
    y = torch.empty(N, H, H, Co, device='cuda'); dy = torch.randn_like(y); dx = torch.empty_like(x)
    dw = torch.zeros_like(w); db = torch.zeros(Co, device='cuda'); b = torch.zeros(Co, device='cuda')
    stats = torch.zeros(2 * Co, dtype=torch.float64, device='cuda')
    tab = (ctypes.c_int64 * 4)(0, k * k, C, Co)
    sp = L.stream_ptr
    L.call('se_split_filters', w.data_ptr(), wt.data_ptr(), wl.data_ptr(), wtl.data_ptr(), tab, 1, sp())
    aux = L.ConvAux(wt.data_ptr(), wtl.data_ptr(), wl.data_ptr())
    out = {}
    for mname in modes:
        mode = MODES[mname]
        fns = {
            'fwd': lambda: L.call('se_conv2d_fwd_aux', d, x.data_ptr(), w.data_ptr(), aux, b.data_ptr(), None, y.data_ptr(), 0, stats.data_ptr(), mode, sp()),
            'dgrad': lambda: L.call('se_conv2d_dgrad_aux', d, dy.data_ptr(), w.data_ptr(), aux, dx.data_ptr(), 0.0, mode, sp()),
            'wgrad': lambda: L.call('se_conv2d_wgrad', d, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), mode, sp()),
        }
        for name, fn in fns.items():
            for _ in range(3): fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps): fn()
            g.replay(); torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); e.record(); torch.cuda.synchronize()
            out['%s_%s' % (name, mname)] = round(1000 * a.elapsed_time(e) / reps, 2)
    return out


if __name__ == '__main__':
    shapes = [(128, 32, 16, 16), (128, 16, 32, 32), (128, 8, 64, 64), (64, 32, 160, 160)]
    if len(sys.argv) > 1 and sys.argv[1] == '1x1':      # ResNet-50 bottleneck 1x1 layers at batch 32
        shapes = [(32, 56, 64, 256, 1), (32, 56, 256, 64, 1), (32, 28, 512, 128, 1), (32, 14, 256, 1024, 1), (32, 14, 1024, 256, 1),
                  (32, 7, 2048, 512, 1)]
    for s in shapes:
        print(json.dumps({'shape': s, 'env': {k: v for k, v in os.environ.items() if k.startswith('SE_')}, 'us': bench(*s, modes=('x3', 'tc', 'f32') if len(s) > 4 else ('x3', 'tc'))}))
