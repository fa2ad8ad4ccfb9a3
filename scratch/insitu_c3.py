"""CTN_TIMING=1 python scratch/insitu_c3.py — per-kernel time of the configs[3] training step (C = 3, 16 x 4 s) in place"""
import os, sys, torch
os.environ["CTN_TIMING"] = "1"
os.environ.setdefault("CTN_NO_PDL", "1")
sys.path.insert(0, "/root/repo")
from conv_tasnet_b200 import ConvTasNet, cal_loss, _lib
from conv_tasnet_b200.optim import FusedAdam
from oracle import conv_tasnet_oracle as O
torch.manual_seed(0)
m = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 3).cuda().train()
opt = FusedAdam(m, lr=1e-3, max_grad_norm=5.0)
mix, src, lens = O.synthetic_batch(16, 32000, 3, 20, 5)
mix, src, lens = mix.cuda(), src.cuda(), lens.cuda()
def step():
    est = m(mix); loss, *_ = cal_loss(src, est, lens); opt.zero_grad(); loss.backward(); opt.step()
for _ in range(2): step()
_lib.lib().ctn_timing_report(1)
for _ in range(3): step()
_lib.lib().ctn_timing_report(0)
