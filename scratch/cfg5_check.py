import sys, time, torch
sys.path.insert(0, '/root/repo')
from conv_tasnet_b200 import ConvTasNet
from oracle import conv_tasnet_oracle as O
torch.manual_seed(0)
model = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 2).cuda().eval()
mix, _, _ = O.synthetic_batch(8, 480000, 2, 20, 1239)
mix = mix.cuda()
with torch.no_grad():
    est = model(mix); torch.cuda.synchronize()
    t = time.time(); est = model(mix); torch.cuda.synchronize(); dt = time.time() - t
    one = model(mix[3:4])
print('cfg5 shape', tuple(est.shape), 'finite', bool(torch.isfinite(est).all()), f'{dt*1e3:.1f} ms for 8 x 60 s ->', 8 * 60 / dt, 'audio-s/s')
print('batch independence err', ((one - est[3:4]).abs().max() / est[3:4].abs().max()).item(), 'peak mem GB', torch.cuda.max_memory_allocated() / 2**30)
# config 3 shape: causal cLN batch 32 x 4 s forward throughput (fp32 I/O)
m3 = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 2, norm_type='cLN', causal=True).cuda().eval()
mix3, _, _ = O.synthetic_batch(32, 32000, 2, 20, 1237); mix3 = mix3.cuda()
with torch.no_grad():
    m3(mix3); torch.cuda.synchronize(); t = time.time()
    for _ in range(5): m3(mix3)
    torch.cuda.synchronize(); dt = (time.time() - t) / 5
print(f'cfg3 (causal cLN, 32 x 4 s, fp32): {dt*1e3:.1f} ms ->', 32 * 4 / dt, 'audio-s/s')
# config 4 shape: C=3 training step batch 16 x 4 s
from conv_tasnet_b200 import cal_loss
from conv_tasnet_b200.optim import FusedAdam
m4 = ConvTasNet(256, 20, 256, 512, 3, 8, 4, 3).cuda().train(); opt = FusedAdam(m4, max_grad_norm=5.0)
mix4, src4, len4 = O.synthetic_batch(16, 32000, 3, 20, 1238); mix4, src4, len4 = mix4.cuda(), src4.cuda(), len4.cuda()
def step():
    est = m4(mix4); loss, *_ = cal_loss(src4, est, len4); opt.zero_grad(); loss.backward(); opt.step(); return loss
for _ in range(2): step()
torch.cuda.synchronize(); t = time.time()
for _ in range(5): l = step()
torch.cuda.synchronize(); dt = (time.time() - t) / 5
print(f'cfg4 (C=3, 16 x 4 s train step, eager): {dt*1e3:.1f} ms ->', 16 * 4 / dt, 'audio-s/s', 'loss', l.item(), 'peak mem GB', torch.cuda.max_memory_allocated() / 2**30)
