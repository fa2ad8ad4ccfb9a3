import sys, torch
sys.path.insert(0,'/root/repo')
from oracle import conv_tasnet_oracle as O
from oracle import fused_schedule as FS
torch.set_num_threads(8)
kw,M,T=dict(C=3),2,12000
cfg=O.Config(**dict(O.PAPER.as_dict(),**kw)); sd=O.init_state_dict(cfg,0)
mix,src,lens=O.synthetic_batch(M,T,cfg.C,cfg.L,1234)
sd64={k:v.double() for k,v in sd.items()}
l64,e64,g64,_,_=O.train_step_grads(cfg,sd64,mix.double(),src.double(),lens)
def er(g): return {k:((g[k].double()-g64[k]).abs().max()/g64[k].abs().max()).item() for k in g64}
def top(c,tag):
    for k in sorted(c,key=lambda k:-c[k])[:3]: print(f'   {tag} {c[k]:.3e} {k}')
def cast(o,dt):
    if isinstance(o,torch.Tensor): return o.to(dt) if o.is_floating_point() else o
    if isinstance(o,dict): return {k:cast(v,dt) for k,v in o.items()}
    if isinstance(o,list): return [cast(v,dt) for v in o]
    return o
est64,st64=FS.model_fwd(cfg,sd64,mix.double())
pf64=FS.pit_fwd(src.double(),est64,lens); d64=FS.pit_bwd(src.double(),pf64['est_masked'],pf64)
est32,st32=FS.model_fwd(cfg,sd,mix)
pf32=FS.pit_fwd(src,est32,lens); d32=FS.pit_bwd(src,pf32['est_masked'],pf32)
print('d_est err fp32 PIT on fp32 est:', ((d32.double()-d64).abs().max()/d64.abs().max()).item())
# fp64 PIT on the fp32 estimate
pfm=FS.pit_fwd(src.double(),est32.double(),lens); dm=FS.pit_bwd(src.double(),pfm['est_masked'],pfm)
print('d_est err fp64 PIT on fp32 est:', ((dm-d64).abs().max()/d64.abs().max()).item())
print('est err', ((est32.double()-est64).abs().max()/est64.abs().max()).item())
top(er(FS.model_bwd(cfg,sd64,mix.double(),st64,d32.double())),'fp64 bwd, fp64 stash, fp32 d_est      ')
top(er(FS.model_bwd(cfg,sd64,mix.double(),st64,dm)),'fp64 bwd, fp64 stash, fp64PIT(fp32 est)')
top(er(FS.model_bwd(cfg,sd64,mix.double(),cast(st32,torch.float64),d64)),'fp64 bwd, fp32 stash, fp64 d_est      ')
# reference-style PIT in fp32 on fp32 est
e=est32.clone().requires_grad_(True); l,_,_,_=O.cal_loss(src,e*1.0,lens); dref,=torch.autograd.grad(l,e)
print('d_est err ref-style fp32 PIT on fp32 est:', ((dref.double()-d64).abs().max()/d64.abs().max()).item())
