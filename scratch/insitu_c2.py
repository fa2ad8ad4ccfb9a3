"""CTN_TIMING=1 python scratch/insitu_c2.py — per-kernel time of the configs[2] forward (causal cLN, 32 x 4 s), bf16 and fp32"""
import os, sys, torch
os.environ["CTN_TIMING"] = "1"
os.environ.setdefault("CTN_NO_PDL", "1")
sys.path.insert(0, "/root/repo")
from conv_tasnet_b200 import ConvTasNet, _lib
from oracle import conv_tasnet_oracle as O
torch.manual_seed(0)
m = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 2, norm_type="cLN", causal=True).cuda().eval()
mix, _, _ = O.synthetic_batch(32, 32000, 2, 20, 5)
mix = mix.cuda()
for half in (True, False):
    m.half_inference(half)
    with torch.no_grad():
        for _ in range(2): m(mix)
        _lib.lib().ctn_timing_report(1)
        for _ in range(3): m(mix)
    print("=== bf16" if half else "=== fp32", flush=True)
    _lib.lib().ctn_timing_report(0)
