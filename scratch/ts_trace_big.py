"""cycle trace of ts_gemm_kernel (library built with -DCTN_TS_TRACE): per-iteration pipeline events of a few CTAs"""
import sys, os, ctypes, torch, numpy as np
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
lib = _lib.lib(); dev = torch.device('cuda:0')
L = ctypes.CDLL(os.environ["CTN_B200_LIB"])
L.ctn_debug_read_ts_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
st = _lib.stream()
def run(F, Kd, O, kn, label):
    A = torch.randn(F, Kd, device=dev); W = (torch.randn(Kd, O, device=dev) if kn else torch.randn(O, Kd, device=dev)) / 16
    D = torch.empty(F, O, device=dev)
    def gemm(): _lib.check(lib.ctn_conv1x1(A.data_ptr(), W.data_ptr(), kn, D.data_ptr(), F, O, Kd, 3199, None, None, None, None, None, None, None, None, st))
    for _ in range(3): gemm()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): gemm()
    e1.record(); torch.cuda.synchronize()
    print(f"=== {label}: F={F} Kd={Kd} O={O} {'bf16' if kn else 'tf32'}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call (incl. split_planes)")
    n = 148
    buf = (ctypes.c_longlong * (256 * n))()
    L.ctn_debug_read_ts_trace(buf, n)
    t = np.frombuffer(buf, dtype=np.int64).reshape(n, 256).astype(np.float64)
    d = t - t[:, 0:1]
    print(" exit (median / max over CTAs):", np.median(d[:, 2]), d[:, 2].max(), " setup done:", np.median(d[:, 1]))
    for cta in (0, 13, 77):
        print(f" CTA {cta}: setup {d[cta,1]:.0f} exit {d[cta,2]:.0f}")
        for it in range(0, 34):
            if t[cta, 8 + it] < t[cta, 0]: break
            print(f"  it{it:2d}: W-prod issue {d[cta,48+it]:7.0f} | conv sees raw {d[cta,88+it]:7.0f} conv done {d[cta,176+it]:7.0f} | mma sees W {d[cta,8+it]:7.0f} A {d[cta,136+it]:7.0f}")
        for s in range(4):
            if t[cta, 216 + s] >= t[cta, 0]: print(f"  epilogue seg{s}: start {d[cta,216+s]:7.0f} end {d[cta,224+s]:7.0f}")
run(383992, 512, 256, 1, "bf16 down big")
run(383992, 256, 512, 1, "bf16 up big")
