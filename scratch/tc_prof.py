import sys, math, torch
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
L = _lib.lib(); dev = torch.device('cuda:0')
F, B, H = 9597, 256, 512
A = torch.randn(F, B, device=dev); W = torch.randn(H, B, device=dev) / 16; D = torch.empty(F, H, device=dev)
G = torch.randn(F, H, device=dev); dW = torch.zeros(H, B, device=dev); W2 = torch.randn(B, H, device=dev) / 22; D2 = torch.empty(F, B, device=dev)
def st(): return _lib.stream()
def gemm_BH(): _lib.check(L.ctn_conv1x1(A.data_ptr(), W.data_ptr(), 0, D.data_ptr(), F, H, B, 3199, None, None, None, None, None, None, None, None, st()))
def gemm_HB(): _lib.check(L.ctn_conv1x1(G.data_ptr(), W2.data_ptr(), 0, D2.data_ptr(), F, B, H, 3199, None, None, None, None, None, None, None, None, st()))
def wgrad(): _lib.check(L.ctn_wgrad(G.data_ptr(), A.data_ptr(), dW.data_ptr(), F, H, B, 3199, None, None, None, None, None, st()))
for fn in (gemm_BH, gemm_HB, wgrad):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    print(f'{fn.__name__}: {us:.2f} us per launch (graph replay, includes the weight-split kernel for gemm), {2*F*B*H/us/1e6:.1f} TFLOP/s')
