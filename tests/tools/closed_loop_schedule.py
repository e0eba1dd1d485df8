"""What the work-queue order is worth in a real closed loop: 4096 episodes (T = 20, two obstacles), device time of loop
iterations 5..44 (all episodes still running) for the three schedules."""
import sys, os, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from junction_mpc.episodes import BatchedEpisodes
from oracle import collision_oracle as C
course = synth.load_course("intersection")
dl = float(np.linalg.norm(course[0,:2]-course[1,:2]))
B, T = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng=np.random.default_rng(11)
state0=np.repeat(np.array([[course[0,0],course[0,1],0.0,course[0,2]]]),B,axis=0); state0[:,2]=rng.uniform(0,3,B)
obst=np.zeros((B,2,6))
k=rng.integers(150,400,(B,2)); ang=rng.uniform(-np.pi,np.pi,(B,2))
obst[:,:,0]=course[k,0]-25*np.cos(ang); obst[:,:,1]=course[k,1]-25*np.sin(ang); obst[:,:,2]=rng.uniform(3,8,(B,2)); obst[:,:,3]=ang; obst[:,:,5]=rng.uniform(-.05,.05,(B,2))
margin=C.cutoff_margin(C.CarGeometry(),dl)
for sched in ("index","apriori","history"):
    engine = BatchedMPC([course], dl=dl, T=T, max_batch=B, schedule=sched)
    ep=BatchedEpisodes(engine,state0,obstacles=obst.copy(),frame_window=10,margin=margin,max_steps=64,record_history=False)
    for _ in range(5): ep.iterate()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40): ep.iterate()
    e1.record(); torch.cuda.synchronize()
    it=ep.out.iters.cpu().numpy()
    print(json.dumps(dict(T=T, schedule=sched, ms_per_loop_iteration=e0.elapsed_time(e1)/40, active=int((ep.done==0).sum().item()), iters_mean=float(it.mean()), iters_max=int(it.max()))), flush=True)
    engine.close()
