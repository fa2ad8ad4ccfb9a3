"""bit-exactness of the CTA-pair GEMM path against the single-CTA path (CTN_NO_PAIR=1 reference saved to disk first)"""
import os, sys, torch
sys.path.insert(0, "/root/repo")
from conv_tasnet_b200 import _lib
lib = _lib.lib(); dev = torch.device("cuda:0")
torch.manual_seed(0)
F, B, H = 9597, 256, 512
A1 = torch.randn(F, B, device=dev); W1 = torch.randn(H, B, device=dev) / 16
A2 = torch.randn(F, H, device=dev); W2 = torch.randn(B, H, device=dev) / 16
st = _lib.stream()
def run(A, W, O, Kd, kn):
    D = torch.empty(F, O, device=dev)
    _lib.check(lib.ctn_conv1x1(A.data_ptr(), W.data_ptr(), kn, D.data_ptr(), F, O, Kd, 3199, None, None, None, None, None, None, None, None, st))
    return D
mode = sys.argv[1]
outs = []
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    outs = [run(A1, W1, H, B, 0), run(A2, W2, B, H, 0), run(A2, W1, B, H, 1), run(A1, W2, H, B, 1)]
    if mode == "check":
        ref = torch.load("/tmp/pair_ref.pt")
        for i, (o, r) in enumerate(zip(outs, ref)):
            if not torch.equal(o.cpu(), r):
                print("MISMATCH iteration", it, "gemm", i, (o.cpu() - r).abs().max().item()); sys.exit(1)
if mode == "save":
    torch.save([o.cpu() for o in outs], "/tmp/pair_ref.pt"); print("saved")
else:
    print("all bit-exact over", sys.argv[2], "iterations x 4 GEMMs")
