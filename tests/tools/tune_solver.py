"""Solver tuning on the GPU box: iteration count and worst parity error against the CPU oracle as a function of the
termination tolerances and the starting point (env knobs JMPC_TOL_RES / JMPC_INIT_MU, option mu_tol)."""
import sys, os, time, json, itertools
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from helpers import oracle_batch, scaled_err
cases=[("config2", synth.make_workload(2), 4096), ("config3", synth.make_workload(3), 3072),
       ("sweep_T20", synth.make_sweep(20, states_per_point=1, max_points=2048), 2048),
       ("sweep_T8", synth.make_sweep(8, states_per_point=1, max_points=2048), 2048)]
refs={}
for name,w,n in cases:
    t0=time.time(); refs[name]=oracle_batch(w, range(n), processes=os.cpu_count()); print(name,"oracle",round(time.time()-t0,1),"s",flush=True)
settings=[dict(mu=1e-13,res=1e-9,init=0)]
for mu,res in [(1e-12,1e-9),(1e-11,1e-8),(1e-10,1e-8),(1e-10,1e-7),(1e-9,1e-7),(1e-8,1e-6)]:
    settings.append(dict(mu=mu,res=res,init=0))
for init in [0.1,1.0,10.0]:
    settings.append(dict(mu=1e-13,res=1e-9,init=init))
report=[]
for st in settings:
    os.environ["JMPC_TOL_RES"]=repr(st["res"]); os.environ["JMPC_INIT_MU"]=repr(st["init"])
    row=dict(st)
    for name,w,n in cases:
        mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=w["B"], mu_tol=st["mu"])
        out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=w["params"])
        worst=0.0; wc=0.0; mism=0
        for k in range(n):
            r=refs[name][k]
            if int(out.status[k])!=r.status: mism+=1; continue
            if r.status!=0: continue
            e=max(scaled_err(out.oa[k],r.oa), scaled_err(out.od[k],r.od), scaled_err(out.ox[k],r.ox), scaled_err(out.oy[k],r.oy), scaled_err(out.ov[k],r.ov), scaled_err(out.oyaw[k],r.oyaw))
            worst=max(worst,e); wc=max(wc, abs(out.cost[k]-r.cost)/abs(r.cost))
        row[name]=dict(iters=round(float(out.iters[:n].mean()),2), imax=int(out.iters[:n].max()), worst=float("%.3g"%worst), cost=float("%.2g"%wc), mism=mism)
        mpc.close()
    report.append(row); print(json.dumps(row),flush=True)
json.dump(report, open('/root/repo/gpurun_out/tune_solver.json','w'), indent=1)
