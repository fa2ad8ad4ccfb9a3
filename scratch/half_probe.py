"""stand-alone check of the single-bf16 (reduced-precision) 1x1-conv flavours, one per process (MODE env)"""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
lib = _lib.lib(); dev = torch.device('cuda:0'); st = _lib.stream()
mode = int(os.environ.get("MODE", 2)); F = int(os.environ.get("F", 2398)); Kd = int(os.environ.get("KD", 256)); O = int(os.environ.get("O", 512))
torch.manual_seed(0)
A = torch.randn(F, Kd, device=dev); W = torch.randn(O, Kd, device=dev) / 16
hi = W.to(torch.bfloat16); lo = (W - hi.float()).to(torch.bfloat16)
Ain = A.to(torch.bfloat16) if mode == 4 else A
D = torch.empty(F, O, device=dev, dtype=torch.bfloat16 if mode == 3 else torch.float32)
print("launch mode", mode, F, Kd, O, flush=True)
_lib.check(lib.ctn_conv1x1_planes(Ain.data_ptr(), hi.data_ptr(), lo.data_ptr(), mode, D.data_ptr(), F, O, Kd, 1199, st))
torch.cuda.synchronize()
Ar = A if mode < 2 else A.to(torch.bfloat16).float()
Wr = W if mode < 2 else hi.float()
want = Ar.double() @ Wr.double().t()
print("mode", mode, "max-rel-err", ((D.double() - want).abs().max() / want.abs().max()).item(), flush=True)
