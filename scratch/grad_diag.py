import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import conv_tasnet_oracle as O
from conv_tasnet_b200 import ConvTasNet, cal_loss
from golden.make_golden_fp64 import sample_index
z = np.load('/root/repo/tests/golden/paper_cfg2_fp64.npz')
cfg = O.PAPER
sd = O.init_state_dict(cfg, seed=int(z['seed_w']))
model = ConvTasNet(**cfg.as_dict()); model.load_state_dict(sd); model = model.cuda().train()
mix, src, lens = O.synthetic_batch(int(z['M']), int(z['T']), cfg.C, cfg.L, int(z['seed_x']))
est = model(mix.cuda()); loss, max_snr, est_m, _ = cal_loss(src.cuda(), est, lens); loss.backward()
sub = est_m.detach().cpu()[..., ::int(z['est_stride'])].double()
print('mode', {k: v for k, v in os.environ.items() if k.startswith('CTN_')}, 'out err', (sub - torch.from_numpy(z['est_sub']).double()).abs().max().item() / float(z['est_abs_max']), 'loss', loss.item(), float(z['loss']))
off = 0; rows = []
for i, (k, p) in enumerate(model.named_parameters()):
    f = p.grad.flatten().cpu().double(); idx = sample_index(f.numel())
    want = torch.from_numpy(z['g_samples'][off:off + len(idx)]).double(); off += len(idx)
    e = ((f[idx] - want).norm() / want.norm()).item()
    rows.append((e / max(float(z['ref32_rel_l2'][i]), 1e-9), e, float(z['ref32_rel_l2'][i]), k, f.numel()))
rows.sort(reverse=True)
print('worst ratio mine/ref32 (rel L2 vs fp64 truth):')
for r in rows[:6]: print(f'  ratio {r[0]:7.2f}  mine {r[1]:.3e}  ref32 {r[2]:.3e}  n={r[4]:7d} {r[3]}')
import statistics
print('median mine', statistics.median(r[1] for r in rows), 'median ref32', statistics.median(r[2] for r in rows))
