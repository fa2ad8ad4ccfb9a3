import sys, torch, time
sys.path.insert(0,'/root/repo')
from oracle import conv_tasnet_oracle as O
from oracle import fused_schedule as FS
torch.set_num_threads(8)
cases=[(dict(mask_nonlinear='softmax'),1,8000),(dict(C=3),2,12000)]
for kw,M,T in cases:
    cfg=O.Config(**dict(O.PAPER.as_dict(),**kw)); sd=O.init_state_dict(cfg,0)
    mix,src,lens=O.synthetic_batch(M,T,cfg.C,cfg.L,1234)
    l32,e32,g32,_,_=O.train_step_grads(cfg,sd,mix,src,lens)
    sd64={k:v.double() for k,v in sd.items()}
    l64,e64,g64,_,_=O.train_step_grads(cfg,sd64,mix.double(),src.double(),lens)
    pf,est,gf=FS.train_step(cfg,sd,mix,src,lens)   # fused schedule, fp32 torch CPU
    def er(g): return {k:((g[k].double()-g64[k]).abs().max()/g64[k].abs().max()).item() for k in g64}
    a,b=er(g32),er(gf)
    print(kw)
    for k in sorted(b,key=lambda k:-b[k])[:8]: print(f'   fused32 {b[k]:.3e}  ref32 {a[k]:.3e}  {k}')
    # mixed: fused backward in fp32 but fed with fp64-accurate d_est?  isolate PIT: use fp64 forward stash cast to fp32
    est64,st64=FS.model_fwd(cfg,sd64,mix.double())
    pf64=FS.pit_fwd(src.double(),est64,lens); d64=FS.pit_bwd(src.double(),pf64['est_masked'],pf64)
    def cast(o):
        if isinstance(o,torch.Tensor): return o.float() if o.is_floating_point() else o
        if isinstance(o,dict): return {k:cast(v) for k,v in o.items()}
        if isinstance(o,list): return [cast(v) for v in o]
        return o
    gb=FS.model_bwd(cfg,sd,mix,cast(st64),d64.float())
    c=er(gb)
    for k in sorted(c,key=lambda k:-c[k])[:4]: print(f'   fp32-backward-only (exact stash, exact d_est) {c[k]:.3e} {k}')
