"""A/B timing of the step kernel for the library named by JMPC_LIB (build variants with -D flags, run them in one
gpurun call): config 3 (65 536 x T=13), a T=8 sweep slice, config 2 (4096 x T=20, cold schedule)."""
import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
dev=torch.device('cuda',0)
t=lambda a,dt: torch.as_tensor(np.ascontiguousarray(a),dtype=dt,device=dev)
for name,w in [("c3_T13_65536", synth.make_workload(3)), ("sweep_T8_65536", synth.make_sweep(8, states_per_point=8)), ("sweep_T25_65536", synth.make_sweep(25, states_per_point=8)), ("c2_T20_4096", synth.make_workload(2))]:
    B,T=w["B"],w["T"]
    mpc=BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=B)
    state,clen,tgt0,oa0,od0=t(w["state"],torch.float64),t(w["course_len"],torch.int32),t(w["target_ind"],torch.int32),t(w["oa"],torch.float64),t(w["od"],torch.float64)
    prm=None if w["params"] is None else t(w["params"],torch.float64)
    out=mpc.alloc_outputs(B); tgt,oa,od=tgt0.clone(),oa0.clone(),od0.clone()
    ts=[]
    for k in range(8):
        tgt.copy_(tgt0); oa.copy_(oa0); od.copy_(od0); mpc.reset_schedule_hints()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); mpc.step(state,tgt,oa,od,out,course_len=clen,params=prm); e1.record(); torch.cuda.synchronize()
        if k>=3: ts.append(e0.elapsed_time(e1))
    print(os.environ.get("JMPC_LIB","default")[-14:], name, "ms %.3f"%np.median(ts), "M solves/s %.2f"%(B/np.median(ts)/1e3), flush=True)
    mpc.close()
