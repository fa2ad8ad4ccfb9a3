import sys, ctypes, torch, numpy as np
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
L = ctypes.CDLL('/root/repo/conv_tasnet_b200/libctn_b200.so'); dev = torch.device('cuda:0')
lib = _lib.lib()
F, B, H = 9597, 256, 512
A = torch.randn(F, B, device=dev); W = torch.randn(H, B, device=dev) / 16; D = torch.empty(F, H, device=dev)
st = _lib.stream()
def gemm(): _lib.check(lib.ctn_conv1x1(A.data_ptr(), W.data_ptr(), 0, D.data_ptr(), F, H, B, 3199, None, None, None, None, None, None, None, None, st))
for _ in range(3): gemm()
torch.cuda.synchronize()
n = 268
buf = (ctypes.c_longlong * (64 * n))()
L.ctn_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
print('rc', L.ctn_debug_read_trace(buf, n))
t = np.frombuffer(buf, dtype=np.int64).reshape(n, 64).astype(np.float64)
base = t[:, 0:1]
d = t - base
def col(i): return d[:, i]
print('per-CTA cycles relative to entry (median over CTAs):')
print(' prologue done      ', np.median(col(1)))
for kb in range(8): print(f' kb{kb}: loads issued {np.median(col(24+kb)):8.0f}  published {np.median(col(40+kb)):8.0f}  mma sees full {np.median(col(8+kb)):8.0f}')
print(' mma all issued     ', np.median(col(2)))
print(' tmem_full seen     ', np.median(col(3)))
print(' epi first issue    ', np.median(col(6)), ' first wait done', np.median(col(7)))
print(' epilogue done      ', np.median(col(4)))
print(' exit               ', np.median(col(5)))
print(' first-wave vs second-wave entry spread (cycles):', np.percentile(t[:,0]-t[:,0].min(), [0, 50, 56, 100]))
