"""CTN_TIMING=1 python scratch/insitu_c4.py — per-kernel time of the configs[4] forward (8 x 60 s, gLN, fp32 I/O) in place"""
import os, sys, torch
os.environ["CTN_TIMING"] = "1"
os.environ.setdefault("CTN_NO_PDL", "1")
sys.path.insert(0, "/root/repo")
from conv_tasnet_b200 import ConvTasNet, _lib
from oracle import conv_tasnet_oracle as O
torch.manual_seed(0)
m = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 2).cuda().eval()
mix, _, _ = O.synthetic_batch(8, 480000, 2, 20, 5)
mix = mix.cuda()
with torch.no_grad():
    for _ in range(2): m(mix)
    _lib.lib().ctn_timing_report(1)
    for _ in range(3): m(mix)
_lib.lib().ctn_timing_report(0)
