import sys, os, ctypes, torch, numpy as np
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
L = ctypes.CDLL(os.environ["CTN_B200_LIB"]); dev = torch.device('cuda:0'); lib = _lib.lib()
L.ctn_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
F, B, H = 9597, 256, 512
for (O, I, label) in ((H, B, "dW1 [512,256]"), (B, H, "dW2 [256,512]")):
    G = torch.randn(F, O, device=dev); X = torch.randn(F, I, device=dev); dW = torch.zeros(O, I, device=dev)
    for _ in range(3): _lib.check(lib.ctn_wgrad(G.data_ptr(), X.data_ptr(), dW.data_ptr(), F, O, I, 3199, None, None, None, None, None, _lib.stream()))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (64 * 512))()
    L.ctn_debug_read_trace(buf, 512)
    t = np.frombuffer(buf, dtype=np.int64).reshape(512, 64).astype(np.float64)
    t = t[:148]; d = t - t[:, 0:1]
    med = lambda c: int(np.median(d[:, c]))
    print(label, ": setup", med(1), "| mma sees full kb0..9:", [med(8 + k) for k in range(10)], "| all MMAs issued", med(2), "| accumulator ready", med(3),
          "| staged", med(6), "| reduce issued + smem read", med(4), "| exit", med(5))
