import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, pickle
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from helpers import oracle_batch, scaled_err, params_from_vector
from oracle import mpc_oracle as O, condensed_model as CM
w = synth.make_sweep(8, states_per_point=1, max_points=4096)
mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=w["B"])
out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=w["params"])
bad = np.nonzero(out.status != 0)[0]
print("bad", bad, out.iters[bad])
print("resid (mu, rp, rd/gscale):", out.record[bad][:, [0,6,7]])
refs = oracle_batch(w, bad, processes=1)
insts=[]
for pos,k in enumerate(bad):
    r=refs[pos]; p=params_from_vector(w["params"][k], 8)
    cq=CM.condense(p, r.xref, r.xbar, w["state"][k], r.reaches_end)
    u,it,ok=CM.ipm_solve(cq)
    print(k, "oracle status", r.status, "gpu err", scaled_err(out.oa[k], r.oa), scaled_err(out.od[k], r.od), "numpy model: it", it, "ok", ok, "err", scaled_err(u[:8], r.oa), "v0", w["state"][k,2], "params", w["params"][k][[0,4,5,6,7,8,9]])
    insts.append((p,r,cq))
pickle.dump(insts, open('/root/repo/gpurun_out/bad_T8.pkl','wb'))
