import sys, torch, time
sys.path.insert(0,'/root/repo')
from oracle import conv_tasnet_oracle as O
torch.set_num_threads(8)
cfg=O.PAPER; sd=O.init_state_dict(cfg,0)
mix,src,lens=O.synthetic_batch(3,32000,2,20,1234)
t=time.time(); l32,e32,g32,_,_=O.train_step_grads(cfg,sd,mix,src,lens); print('fp32',time.time()-t, l32.item())
sd64={k:v.double() for k,v in sd.items()}
t=time.time(); l64,e64,g64,_,_=O.train_step_grads(cfg,sd64,mix.double(),src.double(),lens); print('fp64',time.time()-t, l64.item())
errs=sorted(((( g32[k].double()-g64[k]).abs().max()/g64[k].abs().max()).item(),k) for k in g32)
print('worst fp32-vs-fp64 grad errs:'); [print(f'{e:.3e} {k}') for e,k in errs[-8:]]
print('est err', ((e32.double()-e64).abs().max()/e64.abs().max()).item())
torch.save({k:v.float() for k,v in g64.items()}, '/root/repo/scratch/g64.pt')
