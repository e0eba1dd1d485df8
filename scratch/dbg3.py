import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from helpers import oracle_batch, scaled_err
for name,w in [("sweep25", synth.make_sweep(25, states_per_point=8)), ("config3", synth.make_workload(3))]:
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=w["B"])
    out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=w["params"])
    bad = np.nonzero(out.status != 0)[0]
    print(name, "bad", len(bad), "status values", np.unique(out.status[bad], return_counts=True), "iters", out.iters[bad][:30])
    sel = bad[:12]
    refs = oracle_batch(w, sel, processes=8)
    for pos,k in enumerate(sel):
        r=refs[pos]
        if r.status!=0: print(k, "oracle status", r.status); continue
        print(k, "st", out.status[k], "it", out.iters[k], "oracle", r.status, "err oa %.2e od %.2e ox %.2e ov %.2e cost %.3e"%(scaled_err(out.oa[k],r.oa), scaled_err(out.od[k],r.od), scaled_err(out.ox[k],r.ox), scaled_err(out.ov[k],r.ov), abs(out.cost[k]-r.cost)/abs(r.cost)), "v0 %.6f"%w["state"][k,2])
    print("resid (mu, rp, rd/gscale) of bad:", out.record[bad][:, [0,6,7]])
