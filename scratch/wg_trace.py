import sys, ctypes, torch, numpy as np
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import _lib
L = ctypes.CDLL('/root/repo/conv_tasnet_b200/libctn_b200.so'); dev = torch.device('cuda:0'); lib = _lib.lib()
F, B, H = 9597, 256, 512
G = torch.randn(F, H, device=dev); X = torch.randn(F, B, device=dev); dW = torch.zeros(H, B, device=dev)
for _ in range(3): _lib.check(lib.ctn_wgrad(G.data_ptr(), X.data_ptr(), dW.data_ptr(), F, H, B, 3199, None, None, None, None, None, _lib.stream()))
torch.cuda.synchronize()
n = 148
buf = (ctypes.c_longlong * (64 * 512))()
L.ctn_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.ctn_debug_read_trace(buf, 512)
t = np.frombuffer(buf, dtype=np.int64).reshape(512, 64).astype(np.float64)
# wgrad grid (4,1,37): linear index blockIdx.y*gridDim.x+blockIdx.x collides across z; take rows with plausible data
t = t[:148]
d = t - t[:, 0:1]
print('mma sees full kb0..18:', [int(np.median(d[:, 8 + k])) for k in range(19)])
print('group0 published kb0,2,4..:', [int(np.median(d[:, 30 + k])) for k in range(0, 18, 2)])
print('group1 published kb1,3,5..:', [int(np.median(d[:, 30 + k])) for k in range(1, 18, 2)])
print('mma issued all', np.median(d[:, 2]), 'tmem_full seen', np.median(d[:, 3]), 'epilogue done', np.median(d[:, 4]), 'exit', np.median(d[:, 5]))
