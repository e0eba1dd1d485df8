"""Where the host-path time goes (a -DJMPC_EXPERIMENT build with JMPC_TIMING=1 makes jmpc_step_host_io print its own
breakdown)."""
import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
w = synth.make_workload(2); B, T = w["B"], w["T"]
mpc = BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=B)
hout = mpc.host_outputs(B)
hin = {}
for key in ("state", "target_ind", "oa", "od", "course_len"):
    hin[key] = mpc.pinned_empty(w[key].shape, w[key].dtype); hin[key][...] = w[key]
for mode in ("zero_copy", "zero_copy_results", "staged"):
    mpc.set_host_transfer(mode)
    os.environ.pop("JMPC_TIMING", None)
    for _ in range(3):
        mpc.step_host(hin["state"], hin["target_ind"], hin["oa"], hin["od"], course_len=hin["course_len"], out=hout)
    t0 = time.perf_counter()
    for _ in range(20):
        mpc.step_host(hin["state"], hin["target_ind"], hin["oa"], hin["od"], course_len=hin["course_len"], out=hout)
    print(f"mode {mode}: {(time.perf_counter()-t0)/20*1e3:.3f} ms per call", flush=True)
    os.environ["JMPC_TIMING"] = "1"
    for _ in range(3):
        mpc.step_host(hin["state"], hin["target_ind"], hin["oa"], hin["od"], course_len=hin["course_len"], out=hout)
