"""Throughput of the step kernel and the flag kernel on the larger configs (device-resident inputs)."""
import sys, os, time, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from oracle import collision_oracle as C
dev=torch.device('cuda',0)
t=lambda a,dt: torch.as_tensor(np.ascontiguousarray(a),dtype=dt,device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    ts=[]
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
res={}
for name,w in [("config2_4096xT20", synth.make_workload(2)), ("config3_65536xT13", synth.make_workload(3)), ("config4_262144xT13", synth.make_workload(4, B=int(os.environ.get("B4","262144")))),
               ("sweep_T8_65536", synth.make_sweep(8, states_per_point=8)), ("sweep_T25_65536", synth.make_sweep(25, states_per_point=8))]:
    B,T=w["B"],w["T"]
    mpc=BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=B)
    state,clen,tgt0,oa0,od0=t(w["state"],torch.float64),t(w["course_len"],torch.int32),t(w["target_ind"],torch.int32),t(w["oa"],torch.float64),t(w["od"],torch.float64)
    prm=None if w["params"] is None else t(w["params"],torch.float64)
    out=mpc.alloc_outputs(B)
    tgt,oa,od=tgt0.clone(),oa0.clone(),od0.clone()
    def step():
        tgt.copy_(tgt0); oa.copy_(oa0); od.copy_(od0)
        mpc.step(state,tgt,oa,od,out,course_len=clen,params=prm)
    def restore():
        tgt.copy_(tgt0); oa.copy_(oa0); od.copy_(od0)
    ms=timeit(step)-timeit(restore)
    st=out.status.cpu().numpy(); it=out.iters.cpu().numpy()
    r={"B":B,"T":T,"step_ms":ms,"solves_per_s":B/ms*1e3,"not_optimal":int((st!=0).sum()),"iters_mean":float(it.mean()),"iters_max":int(it.max())}
    if w["obstacles"] is not None:
        geo=C.CarGeometry(); margin=C.cutoff_margin(geo,w["dl"])
        agent,v,obs=t(w["agent_idx"],torch.int32),t(w["state"][:,2],torch.float64),t(w["obstacles"],torch.float64)
        flag=torch.zeros(B,dtype=torch.int32,device=dev); cl=torch.zeros(B,dtype=torch.int32,device=dev)
        cms=timeit(lambda: mpc.collision(agent,v,obs,w["frame_window"],margin,flag,cl))
        r.update(collision_ms=cms, flags_per_s=B/cms*1e3, flag_rate=float(flag.float().mean().item()), n_obs=int(obs.shape[1]))
    res[name]=r
    print(name, json.dumps(r), flush=True)
    del mpc
json.dump(res, open('/root/repo/gpurun_out/configs.json','w'), indent=1)
