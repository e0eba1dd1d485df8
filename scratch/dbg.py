import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from helpers import oracle_batch, scaled_err
w = synth.make_workload(2, B=32)
mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=64)
out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"])
print("status", out.status); print("iters", out.iters)
refs = oracle_batch(w, range(32), processes=8)
for k in range(32):
    r=refs[k]
    print(k, out.status[k], out.iters[k], "tgt", out.target_ind[k], r.target_ind, "xref_eq", np.array_equal(out.xref[k], r.xref),
          "err oa %.2e od %.2e ox %.2e ov %.2e cost %.3e"%(scaled_err(out.oa[k],r.oa), scaled_err(out.od[k],r.od), scaled_err(out.ox[k],r.ox), scaled_err(out.ov[k],r.ov), abs(out.cost[k]-r.cost)/abs(r.cost)))
