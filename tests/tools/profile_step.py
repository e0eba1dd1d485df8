"""A few launches of the step kernel on one configuration, for `ncu --set full -k regex:mpc_step_kernel`:
    python tests/tools/profile_step.py 2|3|T25     (config 2: 4096 x T=20; config 3: 65 536 x T=13; T25: a 65 536 sweep slice)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "av-simulation-at-intersections_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from junction_mpc import synth  # noqa: E402
from junction_mpc.batched import BatchedMPC  # noqa: E402

which = sys.argv[1]
w = synth.make_workload(int(which)) if which in ("2", "3") else synth.make_sweep(25, states_per_point=8)
dev = torch.device("cuda", 0)
t = lambda a, dt: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=w["B"], schedule="apriori")
out = mpc.alloc_outputs(w["B"])
st, clen, prm = t(w["state"], torch.float64), t(w["course_len"], torch.int32), t(w.get("params"), torch.float64)
for _ in range(4):
    mpc.step(st, t(w["target_ind"], torch.int32), t(w["oa"], torch.float64), t(w["od"], torch.float64), out, course_len=clen,
             params=prm)
torch.cuda.synchronize()
print("ok", which, float(out.iters.float().mean()))
