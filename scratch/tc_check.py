"""On-GPU diagnostic of the tcgen05 GEMM / wgrad against fp64 torch, with structured inputs that expose layout bugs."""
import os, sys, math, torch
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
L = _lib.lib(); dev = torch.device('cuda:0')
def P(t): return None if t is None else t.data_ptr()
def conv(A, W, kn=0, **kw):
    F, Kd = A.shape; O = W.shape[1] if kn else W.shape[0]
    D = torch.full((F, O), float('nan'), device=dev)
    _lib.check(L.ctn_conv1x1(P(A), P(W), kn, P(D), F, O, Kd, kw.get('K', F), P(kw.get('alpha_in')), P(kw.get('c1')), P(kw.get('c2')),
                             P(kw.get('acc')), P(kw.get('rs')), P(kw.get('res')), P(kw.get('stat')), P(kw.get('alpha_out')), _lib.stream()))
    torch.cuda.synchronize(); return D
def report(tag, D, want):
    d = (D.double() - want).abs(); e = d.max().item() / want.abs().max().item()
    bad = (d > 1e-3 * want.abs().max()).nonzero()
    print(f'{tag:40s} rel err {e:.3e}  nan {torch.isnan(D).sum().item()}  bad {len(bad)}', flush=True)
    if len(bad): print('   first bad idx', bad[:6].tolist(), 'got', [D[i, j].item() for i, j in bad[:3].tolist()], 'want', [want[i, j].item() for i, j in bad[:3].tolist()])
g = torch.Generator().manual_seed(0)
for (F, Kd, O) in [(144, 64, 128), (288, 128, 128), (9597, 256, 512), (9597, 512, 256), (1000, 256, 768)]:
    print('== shape F,Kd,O', F, Kd, O)
    A = torch.randn(F, Kd, generator=g).to(dev); W = (torch.randn(O, Kd, generator=g) / math.sqrt(Kd)).to(dev)
    ones_w = torch.ones(O, Kd, device=dev); ones_a = torch.ones(F, Kd, device=dev)
    report('A rand, W ones', conv(A, ones_w), A.double() @ ones_w.double().t())
    report('A ones, W rand', conv(ones_a, W), ones_a.double() @ W.double().t())
    oh = torch.zeros(F, Kd, device=dev); oh[torch.arange(F), torch.arange(F) % Kd] = 1
    report('A one-hot(f%Kd), W rand', conv(oh, W), oh.double() @ W.double().t())
    report('A rand, W rand (TN)', conv(A, W), A.double() @ W.double().t())
    report('A rand, W rand (KN)', conv(A, W.t().contiguous(), kn=1), A.double() @ W.double().t())
    res = torch.randn(F, O, generator=g).to(dev); al = torch.tensor([0.3], device=dev)
    stat = torch.zeros(3, 2, dtype=torch.float64, device=dev); K = (F + 2) // 3
    D = conv(A, W, K=K, alpha_in=al, res=res, stat=stat, alpha_out=al)
    want = torch.where(A > 0, A, 0.3 * A).double() @ W.double().t() + res.double()
    report('prelu-in + res', D, want)
    pw = torch.where(want > 0, want, 0.3 * want); m = (torch.arange(F, device=dev) // K)
    ws = torch.stack([torch.stack([pw[m == i].sum(), (pw[m == i] ** 2).sum()]) for i in range(3)])
    print('   stat_out rel err', ((stat - ws).abs().max() / ws.abs().max()).item())
# wgrad
def wgrad(G, X, **kw):
    F, O = G.shape; I = X.shape[1]
    dW = torch.zeros(O, I, device=dev)
    _lib.check(L.ctn_wgrad(P(G), P(X), P(dW), F, O, I, kw.get('K', F), P(kw.get('alpha')), P(kw.get('gamma')), P(kw.get('beta')), P(kw.get('acc')), P(kw.get('rs')), _lib.stream()))
    torch.cuda.synchronize(); return dW
for (F, O, I) in [(64, 128, 128), (256, 128, 256), (9597, 256, 512), (9597, 512, 256), (5000, 768, 256)]:
    print('== wgrad F,O,I', F, O, I)
    G = torch.randn(F, O, generator=g).to(dev); X = torch.randn(F, I, generator=g).to(dev)
    report('G rand, X ones', wgrad(G, torch.ones(F, I, device=dev)), G.double().t() @ torch.ones(F, I, device=dev).double())
    report('G ones, X rand', wgrad(torch.ones(F, O, device=dev), X), torch.ones(F, O, device=dev).double().t() @ X.double())
    report('G rand, X rand', wgrad(G, X), G.double().t() @ X.double())
    gam = torch.randn(I, generator=g).to(dev); bet = torch.randn(I, generator=g).to(dev); al = torch.tensor([0.3], device=dev)
    a = torch.where(X > 0, X, 0.3 * X).double(); mu = a.mean(1, keepdim=True); r = 1 / torch.sqrt(a.var(1, keepdim=True, unbiased=False) + 1e-8)
    rs = torch.cat([mu, r], 1).float().contiguous()
    report('norm prologue (cLN rows)', wgrad(G, X, alpha=al, gamma=gam, beta=bet, rs=rs), G.double().t() @ (gam.double() * (a - mu) * r + bet.double()))
print('done')
