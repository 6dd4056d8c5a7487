"""One launch of every kernel the round-2 ncu --set full capture targets (run under ncu with -k regex:...)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_embeddings_b200 import _lib as L
from semantic_embeddings_b200.evaluate_retrieval import pairwise_distances, pairwise_topk, row_argsort
L.load(); L.check(L.load().se_init())
sp = L.stream_ptr


def conv(N, H, C, Co, mode=2, k=3):
    d = L.ConvDesc(N, H, H, C, Co, k, k, 1, k // 2, k // 2, H, H)
    x = torch.randn(N, H, H, C, device='cuda'); w = torch.randn(k, k, C, Co, device='cuda') * 0.1
    wt, wl, wtl = torch.empty_like(w), torch.empty_like(w), torch.empty_like(w)
    y = torch.empty(N, H, H, Co, device='cuda'); dy = torch.randn_like(y); dx = torch.empty_like(x)
    dw = torch.zeros_like(w); db = torch.zeros(Co, device='cuda'); b = torch.zeros(Co, device='cuda')
    stats = torch.zeros(2 * Co, dtype=torch.float64, device='cuda')
    tab = (ctypes.c_int64 * 4)(0, k * k, C, Co)
    L.call('se_split_filters', w.data_ptr(), wt.data_ptr(), wl.data_ptr(), wtl.data_ptr(), tab, 1, sp())
    aux = L.ConvAux(wt.data_ptr(), wtl.data_ptr(), wl.data_ptr())
    L.call('se_conv2d_fwd_aux', d, x.data_ptr(), w.data_ptr(), aux, b.data_ptr(), None, y.data_ptr(), 0, stats.data_ptr(), mode, sp())
    L.call('se_conv2d_dgrad_aux', d, dy.data_ptr(), w.data_ptr(), aux, dx.data_ptr(), 0.0, mode, sp())
    L.call('se_conv2d_wgrad', d, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), mode, sp())
    torch.cuda.synchronize()


conv(128, 32, 16, 16)        # conv_tc_kernel<1> x2 (fwd, dgrad), conv_wgrad_pk_kernel
conv(128, 16, 32, 32)        # conv_tc_kernel<1> x2, conv_wgrad_tc_kernel<1>
conv(128, 8, 64, 64)         # streamed-weights path, conv_wgrad_tc_kernel<1>
conv(32, 14, 256, 1024, k=1) # ResNet-50 1x1: conv_tc_kernel<1> (flat GEMM) x2, conv1x1_wgrad_tc_kernel<1>
conv(32, 28, 128, 128)       # ResNet-50 3x3 at 28x28: padded row slots, conv_wgrad_tc_kernel<1> with a 32-pixel pitch
n, dd = 50000, 100
f = np.random.RandomState(0).randn(n, dd).astype(np.float32); f /= np.linalg.norm(f, axis=-1, keepdims=True)
fd = torch.from_numpy(f).cuda()
out = torch.empty((n, n), device='cuda')
pairwise_distances(None, False, 0, n, 2, out=out, feat_dev=fd)       # pairwise_tc_kernel<0>
pairwise_topk(k=251, feat_dev=fd)                                      # pairwise_tc_kernel<2>, <1>, pairwise_topk_finish_kernel
row_argsort(out[:296])                                                 # row_argsort_kernel (two CTAs per SM)
torch.cuda.synchronize()
