"""VARIANT=path/to/lib.so python scratch/variant_bench.py — graph-replayed paper-config step time + in-place kernel times
with an alternative build of the library (A/B of compile-time tile constants)."""
import os, sys, torch
os.environ["CTN_TIMING"] = "1" if os.environ.get("INSITU") == "1" else "0"
sys.path.insert(0, "/root/repo")
from conv_tasnet_b200 import _lib
if os.environ.get("VARIANT"):
    _lib._LIB_PATH = os.path.abspath(os.environ["VARIANT"])
from conv_tasnet_b200 import ConvTasNet, cal_loss
from conv_tasnet_b200.optim import FusedAdam
from conv_tasnet_b200.graph import GraphedTrainStep
from oracle import conv_tasnet_oracle as O
torch.manual_seed(0)
m = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 2).cuda().train()
opt = FusedAdam(m, lr=1e-3, max_grad_norm=5.0)
mix, src, lens = O.synthetic_batch(3, 32000, 2, 20, 5)
mix, src, lens = mix.cuda(), src.cuda(), lens.cuda()
if os.environ.get("INSITU") == "1":
    def step():
        est = m(mix); loss, *_ = cal_loss(src, est, lens); opt.zero_grad(); loss.backward(); opt.step()
    for _ in range(3): step()
    _lib.lib().ctn_timing_report(1)
    for _ in range(5): step()
    _lib.lib().ctn_timing_report(0)
else:
    step = GraphedTrainStep(m, opt, warmup=3)
    for _ in range(5): step(mix, src, lens)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): loss = step(mix, src, lens)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    print(os.environ.get("VARIANT", "default"), "ms/step %.3f" % best, "loss %.4f" % loss.item())
