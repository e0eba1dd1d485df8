"""Run a sweep slice, report the instances the solver did not bring to the strict tolerances, dump them."""
import sys, os, pickle
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
T=int(sys.argv[1]) if len(sys.argv)>1 else 25
w=synth.make_sweep(T, states_per_point=32)
mpc=BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=w["B"])
out=mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=w["params"])
bad=np.nonzero(out.status!=0)[0]
print("T",T,"B",w["B"],"not optimal:",bad.tolist(),"iters max",int(out.iters.max()), "hist tail", np.bincount(out.iters)[25:].tolist())
sel=np.concatenate([bad, np.argsort(-out.iters)[:20]])
sel=np.unique(sel)
print("iters of selected", out.iters[sel].tolist())
print("record (debug build: mu, ai, cost, status, target, iters, rp, rd)"); np.set_printoptions(linewidth=200, precision=3)
print(out.record[sel])
pickle.dump(dict(idx=sel, T=T, state=w["state"][sel], target=w["target_ind"][sel], oa=w["oa"][sel], od=w["od"][sel], clen=w["course_len"][sel], params=w["params"][sel], iters=out.iters[sel], status=out.status[sel], dl=w["dl"]), open("/root/repo/gpurun_out/bad_T%d.pkl"%T,"wb"))
