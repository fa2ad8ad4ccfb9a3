"""stand-alone timing of the four 1x1-conv GEMM flavours of the paper step (L2 flushed between launches)"""
import sys, os, torch
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
lib = _lib.lib(); dev = torch.device('cuda:0'); st = _lib.stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(F, Kd, O, kn, label, res=False):
    A = torch.randn(F, Kd, device=dev); W = (torch.randn(Kd, O, device=dev) if kn else torch.randn(O, Kd, device=dev)) / 16
    D = torch.empty(F, O, device=dev); R = torch.randn(F, O, device=dev) if res else None
    def gemm(): _lib.check(lib.ctn_conv1x1(A.data_ptr(), W.data_ptr(), kn, D.data_ptr(), F, O, Kd, 3199, None, None, None, None, None, R.data_ptr() if res else None, None, None, st))
    for _ in range(3): gemm()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gemm(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    want = A.double() @ (W.double() if kn else W.double().t()) + (R.double() if res else 0)
    err = ((D.double() - want).abs().max() / want.abs().max()).item()
    print(f"{label:10s} F={F} Kd={Kd} O={O}: {tot / 10 * 1e3:7.1f} us (incl. split_planes ~4 us)  max-rel-err {err:.2e}", flush=True)
F = int(os.environ.get("F", 9597))
run(F, 256, 512, 0, "tf32 up"); run(F, 512, 256, 0, "tf32 down"); run(F, 256, 512, 1, "bf16 up"); run(F, 512, 256, 1, "bf16 down", res=True)
