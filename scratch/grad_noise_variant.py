"""gradient error of the CUDA path vs fp64 truth (tests/golden/paper_cfg2_fp64.npz) next to the reference's own fp32
error (the figure DESIGN.md section 2 quotes).  CTN_B200_LIB selects the library build."""
import os, sys, statistics, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import load_golden, rel_l2
from golden.make_golden_fp64 import sample_index
from conv_tasnet_b200 import ConvTasNet, cal_loss
from oracle import conv_tasnet_oracle as O
z = load_golden("paper_cfg2_fp64.npz")
cfg = O.PAPER
model = ConvTasNet(**cfg.as_dict()); model.load_state_dict(O.init_state_dict(cfg, seed=int(z["seed_w"]))); model = model.cuda().train()
mix, src, lens = O.synthetic_batch(int(z["M"]), int(z["T"]), cfg.C, cfg.L, int(z["seed_x"]))
est = model(mix.cuda()); loss, max_snr, est_m, _ = cal_loss(src.cuda(), est, lens); loss.backward()
sub = est_m.detach().cpu()[..., ::int(z["est_stride"])].double()
e_out = (sub - torch.from_numpy(z["est_sub"]).double()).abs().max().item() / float(z["est_abs_max"])
off, errs, refs = 0, [], []
for i, (k, p) in enumerate(model.named_parameters()):
    f = p.grad.flatten().cpu().double(); idx = sample_index(f.numel())
    want = torch.from_numpy(z["g_samples"][off:off + len(idx)]).double(); off += len(idx)
    if f.numel() == 1: continue
    errs.append(rel_l2(f[idx], want)); refs.append(float(z["ref32_rel_l2"][i]))
print(os.environ.get("CTN_B200_LIB", "default"), "out err %.2e" % e_out, "loss diff %.2e" % abs(loss.item() - float(z["loss"])),
      "grad median %.2e (ref32 %.2e) ratio %.2f" % (statistics.median(errs), statistics.median(refs), statistics.median(errs) / statistics.median(refs)),
      "max %.2e (ref32 max %.2e)" % (max(errs), max(refs)))
