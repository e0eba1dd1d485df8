import sys, os, time, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from junction_mpc.episodes import BatchedEpisodes
from oracle import collision_oracle as C
course = synth.load_course("intersection")
dl = float(np.linalg.norm(course[0,:2]-course[1,:2]))
B=int(sys.argv[1]) if len(sys.argv)>1 else 4096
T=int(sys.argv[2]) if len(sys.argv)>2 else 13
sched=sys.argv[3] if len(sys.argv)>3 else "history"
use_graph=(sys.argv[4]=="graph") if len(sys.argv)>4 else False
engine = BatchedMPC([course], dl=dl, T=T, max_batch=B, schedule=sched)
rng=np.random.default_rng(11)
state0=np.repeat(np.array([[course[0,0],course[0,1],0.0,course[0,2]]]),B,axis=0); state0[:,2]=rng.uniform(0,3,B)
obst=np.zeros((B,2,6))
k=rng.integers(150,400,(B,2)); ang=rng.uniform(-np.pi,np.pi,(B,2))
obst[:,:,0]=course[k,0]-25*np.cos(ang); obst[:,:,1]=course[k,1]-25*np.sin(ang); obst[:,:,2]=rng.uniform(3,8,(B,2)); obst[:,:,3]=ang; obst[:,:,5]=rng.uniform(-.05,.05,(B,2))
margin=C.cutoff_margin(C.CarGeometry(),dl)
for rep in range(2):
    ep=BatchedEpisodes(engine,state0,obstacles=obst.copy(),frame_window=10,margin=margin,max_steps=400,record_history=(rep==1))
    torch.cuda.synchronize(); t0=time.perf_counter()
    res=ep.run(max_steps=400, use_graph=use_graph)
    dt=time.perf_counter()-t0
print(json.dumps(dict(B=B, T=T, schedule=sched, graph=use_graph, wall_s=dt, iterations=res["iterations"], episodes_per_s=B/dt, done=int((res["done"]==1).sum()), index_rule=int((res["done"]==2).sum()),
      steps_mean=float(res["steps"].mean()), steps_max=int(res["steps"].max()), control_steps_per_s=float(res["steps"].sum()/dt))))
