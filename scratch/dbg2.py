import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from helpers import oracle_batch, scaled_err
T=int(sys.argv[1])
w = synth.make_sweep(T, states_per_point=1, max_points=16)
mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=64)
out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=w["params"])
print("status", out.status); print("iters", out.iters)
refs = oracle_batch(w, range(16), processes=8)
for k in range(16):
    r=refs[k]
    print(k, out.status[k], out.iters[k], "err oa %.2e od %.2e ox %.2e ov %.2e cost %.3e"%(scaled_err(out.oa[k],r.oa), scaled_err(out.od[k],r.od), scaled_err(out.ox[k],r.ox), scaled_err(out.ov[k],r.ov), abs(out.cost[k]-r.cost)/abs(r.cost)))
