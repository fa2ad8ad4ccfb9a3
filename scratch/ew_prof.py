import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conv_tasnet_b200 import _lib
L = _lib.lib(); dev = torch.device('cuda:0')
M, K, H, B = 3, 3199, 512, 256
F = M * K
def P(t): return None if t is None else t.data_ptr()
z1 = torch.randn(M, K, H, device=dev); z2 = torch.empty_like(z1); dz2 = torch.randn_like(z1); dn1 = torch.empty_like(z1)
g1 = torch.randn(H, device=dev); b1 = torch.randn(H, device=dev); Wd = torch.randn(H, 3, device=dev)
a1 = torch.tensor([0.25], device=dev); a2 = torch.tensor([0.3], device=dev)
a = torch.where(z1 > 0, z1, 0.25 * z1).double()
acc = torch.stack([a.sum(dim=(1, 2)), (a * a).sum(dim=(1, 2))], 1).contiguous()
stat = torch.zeros(M, 2, dtype=torch.float64, device=dev); red = torch.zeros(M, 2, dtype=torch.float64, device=dev)
dWd = torch.zeros(H, 3, device=dev); dg = torch.zeros(H, device=dev); db = torch.zeros(H, device=dev); dal = torch.zeros(1, device=dev)
st = lambda: _lib.stream()
tests = {
 'dwconv_fwd +stats': lambda: L.ctn_dwconv_fwd(P(z1), P(a1), P(acc), None, P(g1), P(b1), P(Wd), M, K, H, 3, 4, 0, P(z2), P(stat), P(a2), st()),
 'dwconv_fwd nostat': lambda: L.ctn_dwconv_fwd(P(z1), P(a1), P(acc), None, P(g1), P(b1), P(Wd), M, K, H, 3, 4, 0, P(z2), None, None, st()),
 'dwconv_bwd       ': lambda: L.ctn_dwconv_bwd(P(dz2), P(z1), P(a1), P(acc), None, P(g1), P(b1), P(Wd), M, K, H, 3, 4, 0, P(dn1), P(dWd), P(dg), P(db), P(red), st()),
 'norm_bwd_reduce  ': lambda: L.ctn_norm_bwd_reduce(P(dz2), P(z1), P(a1), P(acc), None, P(g1), M, K, H, P(dg), P(db), P(red), st()),
 'gln_bwd_apply    ': lambda: L.ctn_norm_bwd_apply(P(dn1), P(z1), P(a1), P(acc), None, P(g1), P(red), M, K, H, P(dal), st()),
}
for name, fn in tests.items():
    for _ in range(3): _lib.check(fn())
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): _lib.check(fn())
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1) * 1000 / 20:.2f} us (warm L2, graph replay)')
