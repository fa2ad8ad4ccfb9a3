import sys, ctypes, torch, numpy as np
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import ConvTasNet, cal_loss, _lib
from oracle import conv_tasnet_oracle as O
L = ctypes.CDLL('/root/repo/conv_tasnet_b200/libctn_b200.so')
torch.manual_seed(0)
m = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 2).cuda().train()
mix, src, lens = O.synthetic_batch(3, 32000, 2, 20, 5)
for _ in range(3):
    est = m(mix.cuda()); loss, *_ = cal_loss(src.cuda(), est, lens); m.zero_grad(); loss.backward()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 512))()
L.ctn_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.ctn_debug_read_trace(buf, 512)
t = np.frombuffer(buf, dtype=np.int64).reshape(512, 64).astype(np.float64)[:128]
d = t - t[:, 0:1]
print('mma sees full kb0..39 (median):', [int(np.median(d[:, 8 + k])) for k in range(0, 40, 3)])
print('per-kb period, kb 20..39:', (np.median(d[:, 8 + 39]) - np.median(d[:, 8 + 20])) / 19)
print('last kb seen', np.median(d[:, 6]), 'tmem_full', np.median(d[:, 3]), 'epilogue done', np.median(d[:, 4]))
print('per-CTA last-kb spread (min/med/max):', np.min(d[:, 6]), np.median(d[:, 6]), np.max(d[:, 6]))
