import sys, os, json, pickle
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from junction_mpc.episodes import BatchedEpisodes
from oracle import collision_oracle as C
course = synth.load_course("intersection")
dl = float(np.linalg.norm(course[0,:2]-course[1,:2]))
B, T = 4096, 20
rng=np.random.default_rng(11)
state0=np.repeat(np.array([[course[0,0],course[0,1],0.0,course[0,2]]]),B,axis=0); state0[:,2]=rng.uniform(0,3,B)
obst=np.zeros((B,2,6))
k=rng.integers(150,400,(B,2)); ang=rng.uniform(-np.pi,np.pi,(B,2))
obst[:,:,0]=course[k,0]-25*np.cos(ang); obst[:,:,1]=course[k,1]-25*np.sin(ang); obst[:,:,2]=rng.uniform(3,8,(B,2)); obst[:,:,3]=ang; obst[:,:,5]=rng.uniform(-.05,.05,(B,2))
margin=C.cutoff_margin(C.CarGeometry(),dl)
engine = BatchedMPC([course], dl=dl, T=T, max_batch=B)
ep=BatchedEpisodes(engine,state0,obstacles=obst.copy(),frame_window=10,margin=margin,max_steps=64,record_history=False)
hist=[]
saved=None
for it in range(45):
    pre=dict(state=ep.state.cpu().numpy().copy(), target=ep.target_ind.cpu().numpy().copy(), oa=ep.oa.cpu().numpy().copy(), od=ep.od.cpu().numpy().copy(), warm=ep.warm.cpu().numpy().copy())
    ep.iterate(); torch.cuda.synchronize()
    iters=ep.out.iters.cpu().numpy(); clen=ep.course_len.cpu().numpy()
    hist.append((int(iters.max()), float(iters.mean()), int((iters>=25).sum())))
    if iters.max()>=30 and saved is None:
        sel=np.argsort(-iters)[:8]
        saved=dict(step=it, idx=sel, iters=iters[sel], state=pre["state"][sel], target=pre["target"][sel], oa=pre["oa"][sel], od=pre["od"][sel], warm=pre["warm"][sel], clen=clen[sel], dl=dl)
print("per step (max, mean, count>=25):", hist)
pickle.dump(saved, open('/root/repo/gpurun_out/slow_closed_loop.pkl','wb'))
print("saved", None if saved is None else (saved["step"], saved["iters"].tolist()))
