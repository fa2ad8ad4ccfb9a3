import os, sys, torch
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import ConvTasNet
from oracle import conv_tasnet_oracle as O
norm, causal = os.environ.get("NORM", "cLN"), os.environ.get("CAUSAL", "1") == "1"
cfg = O.Config(**{**O.PAPER.as_dict(), "norm_type": norm, "causal": causal, "R": 1, "X": 2})
model = ConvTasNet(**cfg.as_dict()); model.load_state_dict(O.init_state_dict(cfg, seed=0)); model = model.cuda().eval()
mix, src, lens = O.synthetic_batch(2, 12000, cfg.C, cfg.L, 77)
with torch.no_grad():
    ref = model(mix.cuda()); torch.cuda.synchronize(); print("fp32 ok", flush=True)
    model.half_inference(True)
    got = model(mix.cuda()); torch.cuda.synchronize()
print("bf16 ok, err", ((got - ref).abs().max() / ref.abs().max()).item(), flush=True)
