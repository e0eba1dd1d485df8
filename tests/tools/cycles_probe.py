"""Latency breakdown of the step kernel from a -DJMPC_CYCLES build (JMPC_LIB=build/variants/libjmpc_cycles.so):
one warp alone (B=1) and the full config-2 batch."""
import sys, os, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
names={0:"rows+residuals",1:"tma wait",2:"symv+grad",3:"K assembly+check",4:"cholesky",5:"rhs build (x2)",6:"tri solves (x2)",7:"directions+step (x2)",
       10:"prep",11:"solve total",12:"output",13:"  chol diag",14:"  chol panel",15:"  chol trailing"}
w=synth.make_workload(2); T=w["T"]
def run(sel,label):
    B=len(sel)
    mpc=BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=max(B,1), schedule="index")
    buf=(C.c_uint64*32)()
    mpc.step_host(w["state"][sel], w["target_ind"][sel], w["oa"][sel], w["od"][sel], course_len=w["course_len"][sel])
    mpc._lib.jmpc_debug_cycles(mpc._h, buf, 1)
    out=mpc.step_host(w["state"][sel], w["target_ind"][sel], w["oa"][sel], w["od"][sel], course_len=w["course_len"][sel])
    mpc._lib.jmpc_debug_cycles(mpc._h, buf, 1)
    c=np.array(buf[:],dtype=np.float64); its=float(out.iters.sum())
    tot=c[10]+c[11]+c[12]
    print(f"--- {label}: B={B}, solver iterations {its:.0f}; cycles per instance {tot/B:.0f}; per iteration (solve total / iterations) {c[11]/its:.0f}")
    for k in (10,11,12,0,1,2,3,4,13,14,15,5,6,7):
        per = c[k]/its if k not in (10,12) else c[k]/B
        print(f"  {names[k]:24s} {100*c[k]/tot:5.1f}%   {per:8.0f} cycles per {'iteration' if k not in (10,12) else 'instance'}")
    mpc.close()
k10=6      # an instance of config 2 that takes 10 iterations
run(np.array([k10]),"one warp alone (instance with 10 iterations)")
run(np.arange(4096),"full batch, 16 warps per SM")
