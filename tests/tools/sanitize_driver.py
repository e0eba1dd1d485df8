"""Small invocations of every kernel for compute-sanitizer (tests/tools/sanitize.sh): argv[1] = case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "av-simulation-at-intersections_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from junction_mpc import synth  # noqa: E402
from junction_mpc.batched import BatchedMPC  # noqa: E402

case = sys.argv[1]
if case in ("step_T20", "step_T13", "step_T8", "step_T25", "step_T10"):
    T = int(case.split("T")[1])
    if T == 20:
        w = synth.make_workload(2, B=96)
    elif T == 13:
        w = synth.make_workload(3, B=97)                     # odd: one half-warp without an instance
    else:
        w = synth.make_sweep(T, states_per_point=1, max_points=64) if T in (8, 25) else None
        if w is None:
            rng = np.random.default_rng(T)
            course = synth.load_course("roundabout")
            w = synth.make_states(rng, course, 40, T)
            w.update(T=T, courses=[course], params=None, B=40, dl=float(np.linalg.norm(course[0, :2] - course[1, :2])))
    st = w["state"].copy()
    st[3, 2] = 30.0                                           # one infeasible instance
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=T, max_batch=128, max_T=31)
    out = mpc.step_host(st, w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=w.get("params"))
    print(case, "status counts", np.unique(out.status, return_counts=True))
elif case == "collision":
    w = synth.make_workload(3, B=128)
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=13, max_batch=128)
    flag, clen = mpc.collision_host(w["agent_idx"], w["state"][:, 2], w["obstacles"], frame_window=20, margin=72)
    print(case, "flag rate", flag.mean())
elif case == "episodes":
    from junction_mpc.episodes import BatchedEpisodes, scripted_obstacles
    e = np.load(os.path.join(ROOT, "tests", "golden", "episode_intersection.npz"))
    engine = BatchedMPC([e["course_smoothed"]], dl=float(e["dl"]), T=13, max_batch=8)
    prog = scripted_obstacles([[dict(kind="t_intersection", direction=1, offset=2., turning=False, speed=25 / 3.6, dt=0.2),
                                dict(kind="t_intersection", direction=-1, offset=4., turning=True, speed=25 / 3.6, dt=0.2)]] * 3)
    ep = BatchedEpisodes(engine, np.repeat(e["state"][:1], 3, axis=0), obstacle_program=prog, frame_window=10, margin=72,
                         max_steps=12)
    res = ep.run(max_steps=12, check_every=4)
    print(case, "steps", res["steps"])
elif case == "planner":
    from junction_mpc import planner as P
    z = np.load(os.path.join(ROOT, "tests", "golden", "planner.npz"))
    names = ["intersection_1_1", "roundabout_1_2_big", "multilane_2_3_2_1", "intersection_1_3"]
    g = lambda n, k: z[f"{n}/{k}"]             # noqa: E731
    scenes = [[g(n, "hp")[k, :g(n, "hp_n")[k]] for k in range(len(g(n, "hp_n")))] for n in names]
    pl = P.BatchedPlanner(z["mp_points"], z["mp_total_length"], float(z["car_radius"]), z["car_circle_centers"])
    r = pl.plan(np.stack([g(n, "start") for n in names]), np.stack([g(n, "goal_point") for n in names]),
                np.stack([g(n, "goal_area") for n in names]), [float(g(n, "allowed_dtheta")) for n in names], scenes,
                scene_id=np.arange(4), weights=np.stack([g(n, "weights") for n in names]), max_expansions=512, max_path=32)
    print(case, "status", r.status, "expansions", r.expansions)
else:
    raise SystemExit("unknown case " + case)
