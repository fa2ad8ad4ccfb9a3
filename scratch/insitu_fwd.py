"""CTN_TIMING=1 in-place kernel times of the inference forward (paper config, M x 4 s, M from argv)"""
import os, sys, torch
os.environ["CTN_TIMING"] = "1"; os.environ.setdefault("CTN_NO_PDL", "1")
sys.path.insert(0, "/root/repo")
from conv_tasnet_b200 import ConvTasNet, _lib
from oracle import conv_tasnet_oracle as O
M = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(0)
m = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 2).cuda().eval()
mix, _, _ = O.synthetic_batch(M, 32000, 2, 20, 5); mix = mix.cuda()
with torch.no_grad():
    for _ in range(3): m(mix)
    _lib.lib().ctn_timing_report(1)
    for _ in range(5): m(mix)
    _lib.lib().ctn_timing_report(0)
