"""Does ordering the work queue pay on the half-warp kernels (pairs of similar instances share a warp)?  Needs a
-DJMPC_EXPERIMENT build (JMPC_SCHED_MAX_WAVES is read from the environment there)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "av-simulation-at-intersections_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
dev = torch.device('cuda', 0)
t = lambda a, dt: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
for name, w in [("c3_T13_65536", synth.make_workload(3)), ("c4_T13_262144", synth.make_workload(4)), ("sweep_T8_65536", synth.make_sweep(8, states_per_point=8))]:
    B, T = w["B"], w["T"]
    for mode in ("index", "apriori", "history"):
        mpc = BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=B, schedule=mode)
        state, clen, tgt0, oa0, od0 = t(w["state"], torch.float64), t(w["course_len"], torch.int32), t(w["target_ind"], torch.int32), t(w["oa"], torch.float64), t(w["od"], torch.float64)
        prm = t(w.get("params"), torch.float64)
        out = mpc.alloc_outputs(B); tgt, oa, od = tgt0.clone(), oa0.clone(), od0.clone()
        ts = []
        for k in range(8):
            tgt.copy_(tgt0); oa.copy_(oa0); od.copy_(od0)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); mpc.step(state, tgt, oa, od, out, course_len=clen, params=prm); e1.record(); torch.cuda.synchronize()
            if k >= 3: ts.append(e0.elapsed_time(e1))
        print(os.environ.get("JMPC_SCHED_MAX_WAVES", "16"), name, mode, "ms %.3f" % np.median(ts), "M solves/s %.2f" % (B / np.median(ts) / 1e3), "launches/step", 1 + (mode != "index"), flush=True)
        mpc.close()
