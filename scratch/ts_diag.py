"""forward / gradient error of the paper-config training step against the fp64 golden, for the GEMM variant selected by
the environment (CTN_GEMM_SS, CTN_TS_MASK, CTN_NO_PDL)"""
import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from conv_tasnet_b200 import ConvTasNet, cal_loss
from oracle import conv_tasnet_oracle as O
z = np.load("/root/repo/tests/golden/paper_cfg2_fp64.npz")
cfg = O.PAPER
sd = O.init_state_dict(cfg, seed=int(z["seed_w"]))
model = ConvTasNet(**cfg.as_dict()); model.load_state_dict(sd); model = model.cuda().train()
mix, src, lens = O.synthetic_batch(int(z["M"]), int(z["T"]), cfg.C, cfg.L, int(z["seed_x"]))
outs = []
for rep in range(3):
    est = model(mix.cuda())
    loss, max_snr, est_m, _ = cal_loss(src.cuda(), est, lens)
    sub = est_m.detach().cpu()[..., ::int(z["est_stride"])].double()
    err = (sub - torch.from_numpy(z["est_sub"]).double()).abs().max().item() / float(z["est_abs_max"])
    outs.append(est_m.detach().clone())
    print(f"rep {rep}: fwd max-rel-err {err:.3e}  loss {loss.item():.6f} (golden {float(z['loss']):.6f})", flush=True)
print("run-to-run max diff:", (outs[0] - outs[1]).abs().max().item(), (outs[1] - outs[2]).abs().max().item())
tag = os.environ.get("TAG", "x")
torch.save(outs[0].cpu(), f"/root/repo/gpurun_out/est_{tag}.pt")
