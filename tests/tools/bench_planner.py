"""Throughput of the batched planner (row f4): the recorded scenario / weight variants replicated to a batch, one
kernel launch; beside it the CPU restatement (oracle/planner_oracle.py) on the same searches, all host cores."""
import json
import os
import sys
import time
from multiprocessing import get_context

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "av-simulation-at-intersections_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

with np.load(os.path.join(ROOT, "tests", "golden", "planner.npz")) as _z:        # in memory: the pool's workers are forks
    Z = {k: _z[k] for k in _z.files}
NAMES = [str(n) for n in Z["variants"] if not np.isnan(Z[f"{n}/cost"])]
G = lambda n, k: Z[f"{n}/{k}"]             # noqa: E731


def _cpu(n):
    from oracle import planner_oracle as PO
    pl = PO.Planner(Z["mp_points"], Z["mp_total_length"], float(Z["car_radius"]), Z["car_circle_centers"])
    r = pl.plan(G(n, "start"), G(n, "goal_point"), G(n, "goal_area"), float(G(n, "allowed_dtheta")), G(n, "hp"), G(n, "hp_n"),
                G(n, "weights"))
    return r.cost


def main():
    from junction_mpc import planner as P
    scenes = [[G(n, "hp")[k, :G(n, "hp_n")[k]] for k in range(len(G(n, "hp_n")))] for n in NAMES]
    pl = P.BatchedPlanner(Z["mp_points"], Z["mp_total_length"], float(Z["car_radius"]), Z["car_circle_centers"])
    res = {}
    for B in (len(NAMES), 1024, 16384):
        idx = np.arange(B) % len(NAMES)
        kw = dict(start=np.stack([G(NAMES[i], "start") for i in idx]), goal_point=np.stack([G(NAMES[i], "goal_point") for i in idx]),
                  goal_area=np.stack([G(NAMES[i], "goal_area") for i in idx]),
                  allowed_dtheta=[float(G(NAMES[i], "allowed_dtheta")) for i in idx], scenes=scenes, scene_id=idx,
                  weights=np.stack([G(NAMES[i], "weights") for i in idx]), max_expansions=1024, max_path=32)
        pl.plan(**kw)
        t0 = time.perf_counter()
        r = pl.plan(**kw)
        wall = time.perf_counter() - t0
        assert (r.status == 0).all()
        res[f"B{B}"] = dict(searches=B, kernel_ms=r.kernel_ms, searches_per_s_kernel=B / r.kernel_ms * 1e3,
                            wall_ms=wall * 1e3, searches_per_s_end_to_end=B / wall,
                            expansions_total=int(r.expansions.sum()), expansions_per_s=float(r.expansions.sum()) / r.kernel_ms * 1e3)
        print(B, json.dumps(res[f"B{B}"]), flush=True)
    cores = os.cpu_count() or 1
    with get_context("fork").Pool(cores) as pool:
        pool.map(_cpu, NAMES[:cores])
        t0 = time.perf_counter()
        pool.map(_cpu, NAMES * 2)
        dt = time.perf_counter() - t0
    res["cpu_restatement"] = dict(searches=2 * len(NAMES), cores=cores, searches_per_s=2 * len(NAMES) / dt,
                                  note="oracle/planner_oracle.py (numpy restatement of the reference's search), one search per call")
    print(json.dumps(res["cpu_restatement"]))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(res, open(os.path.join(out, "r2_planner_throughput.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
