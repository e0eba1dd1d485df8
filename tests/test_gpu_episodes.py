"""Device-resident closed-loop episodes against the recorded reference episodes and
against the same loop run on the CPU with the oracle's functions."""
import os

import numpy as np
import pytest

from oracle import collision_oracle as C
from oracle import mpc_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jm():
    import __graft_entry__ as g
    g.build()
    from junction_mpc import synth
    from junction_mpc.batched import BatchedMPC
    from junction_mpc.episodes import BatchedEpisodes
    return synth, BatchedMPC, BatchedEpisodes


# the scripted obstacles of the two reference scenarios (mpc_intersection.py:46-49, mpc_roundabout.py:48-51)
SCENARIO_OBSTACLES = {
    "intersection": [dict(kind="t_intersection", direction=1, offset=2., turning=False, speed=25 / 3.6, dt=0.2),
                     dict(kind="t_intersection", direction=-1, offset=4., turning=True, speed=25 / 3.6, dt=0.2)],
    "roundabout": [dict(kind="roundabout", direction=1, offset=1., turning=True, speed=25 / 3.6, dt=0.2),
                   dict(kind="roundabout", direction=-1, offset=4., turning=True, speed=25 / 3.6, dt=0.2)],
}


@pytest.mark.parametrize("obstacle_source", ["device_program", "recording"])
@pytest.mark.parametrize("name", ["intersection", "roundabout"])
def test_replay_of_reference_episode(jm, golden_dir, name, obstacle_source):
    """Config 1 on the device.  "device_program": nothing comes from the recording but the initial state -- the
    reference's scripted obstacles (moving_obstacles.py) are stepped by a kernel too (obstacle_script=None);
    "recording": their poses are replayed from the fixture, everything else is computed by the kernels."""
    synth, BatchedMPC, BatchedEpisodes = jm
    from junction_mpc.episodes import scripted_obstacles
    e = np.load(os.path.join(golden_dir, f"episode_{name}.npz"))
    course = e["course_smoothed"]
    n = len(e["state"])
    engine = BatchedMPC([course], dl=float(e["dl"]), T=13, max_batch=8)
    state0 = np.repeat(e["state"][:1], 3, axis=0)                      # three identical egos
    if obstacle_source == "recording":
        kw = dict(obstacle_script=np.repeat(e["obs"][:, None], 3, axis=1))      # [steps, B, n_obs, 6]
    else:
        kw = dict(obstacle_script=None, obstacle_program=scripted_obstacles([SCENARIO_OBSTACLES[name]] * 3))
    ep = BatchedEpisodes(engine, state0, frame_window=int(e["frame_window"]), margin=int(e["margin"]), max_steps=160, **kw)
    res = ep.run(check_every=4)
    assert (res["done"] == 1).all()
    assert (res["steps"] == n).all()                                    # 91 / 116 iterations, as the reference
    h = res["history"][:n, 0]
    # history row i holds the state AFTER step i
    states_after = np.vstack([e["state"][1:], e["final_state"][None]])
    np.testing.assert_allclose(h[:, [0, 1, 3, 2]], states_after, rtol=0, atol=1e-5)
    np.testing.assert_allclose(h[:, 5], e["di"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(h[:, 6], e["ai"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(h[:, 7], e["dev"], rtol=0, atol=1e-5)
    assert np.array_equal(res["flags"][:n, 0], e["flag"])
    np.testing.assert_allclose(res["state"][0], e["final_state"], rtol=0, atol=1e-5)
    assert np.array_equal(res["history"][:n, 1], res["history"][:n, 0])  # identical egos stay identical


def _cpu_episode(course, p, state0, obstacles, fw, margin, max_steps):
    """The scenario loop with constant-input obstacles, on the CPU with the oracle's functions."""
    geo = C.CarGeometry()
    st = tuple(state0)
    obs = [list(o) for o in obstacles]
    agent, tmp_len, target, oa, od, di = 0, None, 0, None, None, 0.0
    n_full = len(course)
    log = []
    for i in range(max_steps):
        d = np.hypot(st[0] - course[-1, 0], st[1] - course[-1, 1])
        cur_len = n_full if tmp_len is None else tmp_len
        if d <= p.goal_dis and abs(target - cur_len) < 5 and abs(st[2]) <= p.stop_speed:
            break
        if tmp_len is None or np.any(course[min(agent, tmp_len - 1)] != course[tmp_len - 1]):
            agent = O.nearest_index_forward(st[0], st[1], course[:, 0], course[:, 1], agent)
        flag, cut = C.collision_cut(geo, course, agent, st[2], obs, dt=p.dt, frame_window=fw, max_accel=p.max_accel,
                                    max_speed=p.sim_max_speed, margin=margin)
        tmp_len = cut
        r = O.mpc_step(p, st, oa, od, course[:cut, 0], course[:cut, 1], course[:cut, 2], target)
        target = r.target_ind
        if r.oa is not None:
            oa, od, di, ai = r.oa, r.od, float(r.od[0]), float(r.oa[0])
        else:
            oa = od = None
            ai = p.max_decel
        st = O.plant_step(p, st, ai, di)
        log.append(st + (int(flag),))
        for o in obs:        # constant-input motion, as their predictor does it
            o[0] += o[2] * np.cos(o[3]) * p.dt
            o[1] += o[2] * np.sin(o[3]) * p.dt
            o[2] += o[4] * p.dt
            o[3] += (o[2] / p.L) * np.tan(o[5]) * p.dt
    return np.array(log)


def test_synthetic_episodes_match_cpu_loop(jm):
    synth, BatchedMPC, BatchedEpisodes = jm
    from helpers import params_from_vector
    course = synth.load_course("intersection")
    dl = float(np.linalg.norm(course[0, :2] - course[1, :2]))
    engine = BatchedMPC([course], dl=dl, T=13, max_batch=16)
    p = params_from_vector(engine.default_params, 13)
    rng = np.random.default_rng(7)
    B, steps = 6, 40
    state0 = np.repeat(np.array([[course[0, 0], course[0, 1], 0.0, course[0, 2]]]), B, axis=0)
    state0[:, 2] = rng.uniform(0, 3, B)
    obstacles = np.zeros((B, 2, 6))
    for b in range(B):
        for o in range(2):
            k = int(rng.integers(150, 400))
            ang = rng.uniform(-np.pi, np.pi)
            obstacles[b, o] = (course[k, 0] - 25 * np.cos(ang), course[k, 1] - 25 * np.sin(ang), rng.uniform(3, 8), ang,
                               0.0, rng.uniform(-0.05, 0.05))
    geo = C.CarGeometry()
    margin = C.cutoff_margin(geo, dl)
    ep = BatchedEpisodes(engine, state0, obstacles=obstacles.copy(), frame_window=10, margin=margin, max_steps=steps)
    res = ep.run(max_steps=steps)
    flagged = 0
    for b in range(B):
        ref = _cpu_episode(course, p, state0[b], obstacles[b], 10, margin, steps)
        got = res["history"][:len(ref), b]
        np.testing.assert_allclose(got[:, [0, 1, 3, 2]], ref[:, :4], rtol=0, atol=1e-5)
        assert np.array_equal(res["flags"][:len(ref), b], ref[:, 4].astype(np.int32))
        flagged += int(ref[:, 4].sum())
    assert flagged > 0          # the obstacles do interfere in this scene


def test_graph_replay_equals_plain_launches(jm):
    """The loop body replayed as a CUDA graph (8 iterations per launch) and launched kernel by kernel give bit-identical
    histories, flags and step counts."""
    synth, BatchedMPC, BatchedEpisodes = jm
    course = synth.load_course("intersection")
    dl = float(np.linalg.norm(course[0, :2] - course[1, :2]))
    rng = np.random.default_rng(3)
    B = 96
    state0 = np.repeat(np.array([[course[0, 0], course[0, 1], 0.0, course[0, 2]]]), B, axis=0)
    state0[:, 2] = rng.uniform(0, 3, B)
    k = rng.integers(150, 400, (B, 2))
    ang = rng.uniform(-np.pi, np.pi, (B, 2))
    obst = np.zeros((B, 2, 6))
    obst[:, :, 0] = course[k, 0] - 25 * np.cos(ang)
    obst[:, :, 1] = course[k, 1] - 25 * np.sin(ang)
    obst[:, :, 2] = rng.uniform(3, 8, (B, 2))
    obst[:, :, 3] = ang
    obst[:, :, 5] = rng.uniform(-.05, .05, (B, 2))
    margin = C.cutoff_margin(C.CarGeometry(), dl)
    out = []
    for use_graph in (False, True):
        engine = BatchedMPC([course], dl=dl, T=13, max_batch=B)
        ep = BatchedEpisodes(engine, state0, obstacles=obst.copy(), frame_window=10, margin=margin, max_steps=400)
        out.append(ep.run(max_steps=400, use_graph=use_graph))
        engine.close()
    a, b = out
    assert (a["done"] == 1).all() and np.array_equal(a["done"], b["done"])
    assert np.array_equal(a["steps"], b["steps"]) and np.array_equal(a["state"], b["state"])
    n = int(a["steps"].max())
    assert np.array_equal(a["history"][:n], b["history"][:n], equal_nan=True)
    assert np.array_equal(a["flags"][:n], b["flags"][:n])


def test_scripted_obstacles_match_the_reference_tracks(jm, golden_dir):
    """The reference's MovingObstacleTIntersection / Roundabout / Arterial, 51 parameter combinations x 160 steps
    recorded from its own classes: the device kernel reproduces every get() tuple."""
    import ctypes as Ct
    import torch
    synth, BatchedMPC, BatchedEpisodes = jm
    from junction_mpc import _cabi
    from junction_mpc.episodes import scripted_obstacles
    z = np.load(os.path.join(golden_dir, "scripted_obstacles.npz"))
    specs, tracks = z["specs"], z["tracks"]
    kinds = {1: "t_intersection", 2: "roundabout", 3: "arterial"}
    rows = [[dict(kind=kinds[int(s[0])], direction=int(s[1]), turning=bool(s[2]), speed=float(s[3]),
                  offset=None if s[4] < 0 else float(s[4]), dt=float(s[5]), x_init=float(s[6]), y_init=float(s[7]),
                  initial_speed=float(s[8]))] for s in specs]
    script, model = scripted_obstacles(rows)
    B = len(rows)
    engine = BatchedMPC([synth.load_course("intersection")], dl=0.083, T=13, max_batch=64)
    dev = torch.device("cuda", 0)
    d_script, d_model = torch.as_tensor(script, device=dev), torch.as_tensor(model, device=dev)
    d_obs = torch.zeros(B, 1, 6, dtype=torch.float64, device=dev)
    p = lambda t: Ct.c_void_p(t.data_ptr())        # noqa: E731
    got = []
    for i in range(tracks.shape[1]):
        _cabi.check(engine._lib.jmpc_scripted_obstacle_step(engine._h, B, 1, p(d_script), p(d_model), p(d_obs), None,
                                                            0 if i == 0 else 1, None), "jmpc_scripted_obstacle_step")
        got.append(d_obs[:, 0].cpu().numpy().copy())
    got = np.stack(got, axis=1)                     # [B, steps, 6]
    # discrete outputs (speed on / off, steering rule) exact; poses to rounding of sin / cos / tan
    assert np.array_equal(got[:, :, 2], tracks[:, :, 2]) and np.array_equal(got[:, :, 5], tracks[:, :, 5])
    np.testing.assert_allclose(got, tracks, rtol=0, atol=1e-9)
