import sys, torch, time
sys.path.insert(0,'/root/repo')
from oracle import conv_tasnet_oracle as O
torch.set_num_threads(8)
cases=[(dict(C=3),2,12000),(dict(norm_type='cLN',causal=True),2,12000),(dict(mask_nonlinear='softmax'),1,8000)]
for kw,M,T in cases:
    cfg=O.Config(**dict(O.PAPER.as_dict(),**kw)); sd=O.init_state_dict(cfg,0)
    mix,src,lens=O.synthetic_batch(M,T,cfg.C,cfg.L,1234)
    l32,e32,g32,_,_=O.train_step_grads(cfg,sd,mix,src,lens)
    l64,e64,g64,_,_=O.train_step_grads(cfg,{k:v.double() for k,v in sd.items()},mix.double(),src.double(),lens)
    errs=sorted((((g32[k].double()-g64[k]).abs().max()/g64[k].abs().max()).item(),k) for k in g32)
    print(kw,'loss',l64.item(),'est err',((e32.double()-e64).abs().max()/e64.abs().max()).item())
    for e,k in errs[-4:]: print(f'   {e:.3e} {k}  |g|max={g64[k].abs().max().item():.3e}')
