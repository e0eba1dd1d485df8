"""One-off large parity validation against the CPU oracle (uses all host cores of the GPU box)."""
import sys, os, time, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av-simulation-at-intersections_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
from helpers import oracle_batch, scaled_err
report={}
big = len(sys.argv) > 1 and sys.argv[1] == "big"
k = 4 if big else 1
cases=[("config2_full", synth.make_workload(2), None), ("config2_other_seed", synth.make_workload(2, seed_offset=5), None),
       (f"config3_first{8192*k}", synth.make_workload(3), 8192 * k),
       (f"config4_first{4096*k}", synth.make_workload(4, B=4096 * k), 4096 * k),
       (f"sweep_T8_{4096*k}pts", synth.make_sweep(8, states_per_point=1, max_points=4096 * k), None),
       (f"sweep_T13_{4096*k}pts", synth.make_sweep(13, states_per_point=1, max_points=4096 * k), None),
       (f"sweep_T20_{2048*k}pts", synth.make_sweep(20, states_per_point=1, max_points=2048 * k), None),
       (f"sweep_T25_{2048*k}pts", synth.make_sweep(25, states_per_point=1, max_points=2048 * k), None)]
for name,w,limit in cases:
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=w["B"])
    out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=w["params"])
    n = limit or w["B"]
    t0=time.time(); refs = oracle_batch(w, range(n), processes=os.cpu_count()); dt=time.time()-t0
    worst=0.0; worst_cost=0.0; mism=0; exact=0
    for k in range(n):
        r=refs[k]
        if int(out.status[k])!=r.status: mism+=1; continue
        if r.status!=0: continue
        exact += int(out.target_ind[k]==r.target_ind and np.array_equal(out.xref[k], r.xref))
        e=max(scaled_err(out.oa[k],r.oa), scaled_err(out.od[k],r.od), scaled_err(out.ox[k],r.ox), scaled_err(out.oy[k],r.oy), scaled_err(out.ov[k],r.ov), scaled_err(out.oyaw[k],r.oyaw))
        worst=max(worst,e); worst_cost=max(worst_cost, abs(out.cost[k]-r.cost)/abs(r.cost))
    report[name]=dict(instances=n, T=w["T"], status_mismatches=mism, index_and_xref_exact=exact, worst_scaled_error=worst, worst_cost_rel=worst_cost,
                      gpu_not_optimal=int((out.status[:n]!=0).sum()), iters_mean=float(out.iters[:n].mean()), iters_max=int(out.iters[:n].max()), oracle_seconds=round(dt,1))
    print(name, json.dumps(report[name]), flush=True)
report['total_instances']=sum(v['instances'] for v in report.values())
print('total', report['total_instances'])
json.dump(report, open('/root/repo/gpurun_out/validate_big.json','w'), indent=1)
