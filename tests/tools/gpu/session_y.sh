set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r2y_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r2y_smoke.log
python - <<'PY' 2>&1 | tee gpurun_out/r2y_latency.log
import sys, time
sys.path.insert(0, 'av-simulation-at-intersections_b200'); sys.path.insert(0, 'tests')
import numpy as np
from junction_mpc import synth
from junction_mpc.batched import BatchedMPC
for cfg, T in ((2, 20), (3, 13), (5, 8), (5, 25)):
    w = synth.make_workload(cfg, B=256) if cfg != 5 else synth.make_sweep_sample(T, 256)
    for label, kw in (("low-latency", {}), ("throughput kernel", {"warps_per_sm": 16})):
        mpc = BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=256, **kw)
        res = {}
        for B in (1, 8, 64, 256):
            sel = slice(0, B)
            prm = None if w["params"] is None else w["params"][sel]
            ts = []
            for k in range(40):
                t0 = time.perf_counter()
                mpc.step_host(w["state"][sel], w["target_ind"][sel], w["oa"][sel], w["od"][sel], course_len=w["course_len"][sel], params=prm)
                ts.append((time.perf_counter() - t0) * 1e3)
            res[B] = float(np.percentile(ts[10:], 50))
        print(f"T={T:2d} {label:18s} step_host p50 ms:", {b: round(v, 3) for b, v in res.items()}, flush=True)
        mpc.close()
PY
