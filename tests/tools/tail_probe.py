"""How much of the step time is scheduling tail?  Config 2 in index order, oracle longest-first order, and the
library's own schedules (a-priori key, previous-step iteration counts)."""
import sys, os, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
dev=torch.device('cuda',0)
t=lambda a,dt: torch.as_tensor(np.ascontiguousarray(a),dtype=dt,device=dev)
cfg=int(sys.argv[1]) if len(sys.argv)>1 else 2
seed=int(sys.argv[2]) if len(sys.argv)>2 else 0
cfgB=int(sys.argv[3]) if len(sys.argv)>3 else 0
w=(synth.make_workload(cfg,B=cfgB or None,seed_offset=seed) if cfg<10 else synth.make_workload(2,B=cfg,seed_offset=seed)); B,T=w["B"],w["T"]
mpc=BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=B, schedule="index")
def run(order, reps=10, reset=False):
    state,clen,tgt0,oa0,od0=t(w["state"][order],torch.float64),t(w["course_len"][order],torch.int32),t(w["target_ind"][order],torch.int32),t(w["oa"][order],torch.float64),t(w["od"][order],torch.float64)
    out=mpc.alloc_outputs(B); tgt,oa,od=tgt0.clone(),oa0.clone(),od0.clone(); ts=[]
    for k in range(reps):
        tgt.copy_(tgt0); oa.copy_(oa0); od.copy_(od0)
        if reset: mpc.reset_schedule_hints()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); mpc.step(state,tgt,oa,od,out,course_len=clen); e1.record(); torch.cuda.synchronize()
        if k>=3: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out.iters.cpu().numpy()
ident=np.arange(B)
ms0,it=run(ident)
print("B=%d T=%d index order  %.3f ms; iters mean %.2f max %d"%(B,T,ms0,it.mean(),it.max()))
lpt=np.argsort(-it,kind="stable")
ms1,_=run(lpt); print("oracle longest first  %.3f ms"%ms1)
mpc.set_schedule("apriori"); ms,_=run(ident); print("schedule=apriori      %.3f ms"%ms)
mpc.set_schedule("history"); ms,_=run(ident,reset=True); print("schedule=history cold %.3f ms (hints reset before every step)"%ms)
ms,it2=run(ident); print("schedule=history warm %.3f ms"%ms, "same results", bool((it2==it).all()))
