import os, sys, torch, statistics
sys.path.insert(0, '/root/repo')
from oracle import conv_tasnet_oracle as O
from conv_tasnet_b200 import ConvTasNet, cal_loss
torch.set_num_threads(os.cpu_count())
kw = dict(mask_nonlinear='softmax'); M, T = 2, 12000
cfg = O.Config(**dict(O.PAPER.as_dict(), **kw)); sd = O.init_state_dict(cfg, 0)
mix, src, lens = O.synthetic_batch(M, T, cfg.C, cfg.L, 1234)
cache = '/tmp/g64_softmax.pt'
if os.path.exists(cache): g64, g32, e64 = torch.load(cache)
else:
    l64, e64, g64, _, _ = O.train_step_grads(cfg, {k: v.double() for k, v in sd.items()}, mix.double(), src.double(), lens)
    l32, e32, g32, _, _ = O.train_step_grads(cfg, sd, mix, src, lens)
    torch.save((g64, g32, e64), cache)
def run(tag):
    model = ConvTasNet(**cfg.as_dict()); model.load_state_dict(sd); model = model.cuda().train()
    est = model(mix.cuda()); loss, _, est_m, _ = cal_loss(src.cuda(), est, lens); loss.backward()
    oe = ((est_m.detach().cpu().double() - e64).abs().max() / e64.abs().max()).item()
    rows = []
    for k, p in model.named_parameters():
        if p.numel() == 1: continue
        w = g64[k].double().flatten(); e = ((p.grad.cpu().double().flatten() - w).norm() / w.norm()).item()
        r = ((g32[k].double().flatten() - w).norm() / w.norm()).item()
        rows.append((e, r, k))
    rows.sort(reverse=True)
    print(tag, 'out err', oe, 'median mine', statistics.median(r[0] for r in rows), 'median ref32', statistics.median(r[1] for r in rows))
    for r in rows[:5]: print(f'   mine {r[0]:.3e} ref32 {r[1]:.3e} {r[2]}')
run(str({k: v for k, v in os.environ.items() if k.startswith('CTN_')}))
