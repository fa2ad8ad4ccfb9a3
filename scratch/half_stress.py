"""repeated launches of the single-bf16 1x1-conv flavours at several sizes, each checked against fp64"""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
lib = _lib.lib(); dev = torch.device('cuda:0'); st = _lib.stream()
torch.manual_seed(0)
bad = 0
for F in (2398, 9597, 51184, 102368, 383992):
    for (mode, Kd, O) in ((2, 256, 256), (3, 256, 512), (4, 512, 256), (4, 256, 512), (0, 256, 512), (0, 512, 256)):
        A = torch.randn(F, Kd, device=dev); W = torch.randn(O, Kd, device=dev) / 16
        hi = W.to(torch.bfloat16); lo = (W - hi.float()).to(torch.bfloat16)
        Ain = A.to(torch.bfloat16) if mode == 4 else A
        D = torch.empty(F, O, device=dev, dtype=torch.bfloat16 if mode == 3 else torch.float32)
        Ar = A if mode < 2 else A.to(torch.bfloat16).float()
        Wr = W if mode < 2 else hi.float()
        want = Ar @ Wr.t()
        worst = 0.0
        for rep in range(20):
            D.zero_()
            _lib.check(lib.ctn_conv1x1_planes(Ain.data_ptr(), hi.data_ptr(), lo.data_ptr(), mode, D.data_ptr(), F, O, Kd, 3199, st))
            worst = max(worst, ((D.float() - want).abs().max() / want.abs().max()).item())
        tol = 4e-3 if mode == 3 else 2e-5
        flag = "" if worst < tol else "  <-- BAD"
        bad += worst >= tol
        print(f"F={F:6d} mode {mode} {Kd}->{O}: worst of 20 launches {worst:.2e}{flag}", flush=True)
print("BAD" if bad else "ALL OK")
