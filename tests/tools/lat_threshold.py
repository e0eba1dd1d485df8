"""Where does the low-latency kernel stop paying?  Device time of one step for batches of 1 .. 16 warps per SM with the
low-latency path forced on / off (-DJMPC_EXPERIMENT build: JMPC_LAT_WARPS_PER_SM is read from the environment)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "av-simulation-at-intersections_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from junction_mpc import synth  # noqa: E402
from junction_mpc.batched import BatchedMPC  # noqa: E402

dev = torch.device("cuda", 0)
t = lambda a, dt: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
for cfg, T in ((2, 20), (3, 13), (5, 8), (5, 25)):
    groups = 2 if T <= 15 else 1
    w = synth.make_workload(cfg, B=148 * 16 * groups) if cfg != 5 else synth.make_sweep_sample(T, 148 * 16 * groups)
    for warps in (1, 2, 3, 4, 6, 8, 12, 16):
        B = 148 * warps * groups
        row = []
        for lat in ("0", "1000"):
            os.environ["JMPC_LAT_WARPS_PER_SM"] = lat
            mpc = BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=B, schedule="index")
            st, tg0, oa0, od0 = t(w["state"][:B], torch.float64), t(w["target_ind"][:B], torch.int32), t(w["oa"][:B], torch.float64), t(w["od"][:B], torch.float64)
            cl, prm = t(w["course_len"][:B], torch.int32), t(None if w["params"] is None else w["params"][:B], torch.float64)
            out = mpc.alloc_outputs(B)
            ts = []
            for k in range(8):
                tg, oa, od = tg0.clone(), oa0.clone(), od0.clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); mpc.step(st, tg, oa, od, out, course_len=cl, params=prm); e1.record(); torch.cuda.synchronize()
                if k >= 3:
                    ts.append(e0.elapsed_time(e1))
            row.append(float(np.median(ts)))
            mpc.close()
        print(f"T={T:2d} {warps:2d} warps per SM (B={B:5d}): throughput kernel {row[0]:.3f} ms, low-latency kernel {row[1]:.3f} ms", flush=True)
