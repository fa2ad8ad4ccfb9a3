"""fp64 truth for the full-size training step (BASELINE configs[1]: paper config, M=3 x 4 s), from the oracle.

    python tests/golden/make_golden_fp64.py        (~1 min on 8 cores; needs no reference checkout)

Why fp64: the reference's own fp32 autograd is up to 2.6e-3 away from fp64 on the PReLU-slope gradients of this config
(4.9 M-term sums), i.e. noisier than the 1e-3 gradient tolerance, so the tolerance is applied against fp64 truth.  The
fp32 oracle is pinned to the reference by tests/test_oracle_golden.py; this file is the same code in double precision.
Stored: per-parameter gradient norm + 64 evenly spaced entries, sub-sampled output, loss."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import conv_tasnet_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count())
cfg = O.PAPER
M, T, SEED_W, SEED_X, NS = 3, 32000, 0, 1234, 64
sd = O.init_state_dict(cfg, SEED_W)
mix, src, lens = O.synthetic_batch(M, T, cfg.C, cfg.L, SEED_X)
sd64 = {k: v.double() for k, v in sd.items()}
loss, est, grads, max_snr, reord = O.train_step_grads(cfg, sd64, mix.double(), src.double(), lens)
arrays = dict(M=np.int64(M), T=np.int64(T), seed_w=np.int64(SEED_W), seed_x=np.int64(SEED_X), loss=np.float64(loss.item()),
              max_snr=max_snr.numpy(), est_stride=np.int64(37), est_sub=est.numpy()[..., ::37].astype(np.float32),
              est_abs_max=np.float64(est.abs().max().item()), names=np.array(list(grads.keys())))
gn, gmax, samp = [], [], []
for k, g in grads.items():
    f = g.flatten()
    idx = torch.linspace(0, f.numel() - 1, NS).long()
    gn.append(f.norm().item())
    gmax.append(f.abs().max().item())
    samp.append(f[idx].numpy())
arrays.update(g_norm=np.array(gn), g_absmax=np.array(gmax), g_samples=np.stack(samp))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "paper_cfg2_fp64.npz")
np.savez_compressed(out, **arrays)
print(out, os.path.getsize(out) // 1024, "KiB")
