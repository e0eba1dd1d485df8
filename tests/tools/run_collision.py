import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from oracle import collision_oracle as C
w=synth.make_workload(3)
B=w["B"]; dev=torch.device('cuda',0)
t=lambda a,dt: torch.as_tensor(np.ascontiguousarray(a),dtype=dt,device=dev)
mpc=BatchedMPC(w["courses"], dl=w["dl"], T=13, max_batch=B)
margin=C.cutoff_margin(C.CarGeometry(),w["dl"])
agent,v,obs=t(w["agent_idx"],torch.int32),t(w["state"][:,2],torch.float64),t(w["obstacles"],torch.float64)
flag=torch.zeros(B,dtype=torch.int32,device=dev); cl=torch.zeros(B,dtype=torch.int32,device=dev)
for _ in range(3): mpc.collision(agent,v,obs,w["frame_window"],margin,flag,cl)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); mpc.collision(agent,v,obs,w["frame_window"],margin,flag,cl); e1.record(); torch.cuda.synchronize()
print("collision ms", e0.elapsed_time(e1), "flag rate", flag.float().mean().item())
