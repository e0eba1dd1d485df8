"""The oracle against the fixtures generated from the UNMODIFIED reference (tests/golden/make_golden.py).

Rows 2-7 and 9-12 of SURVEY.md section 8(a): integers exact, floats to 1e-12."""
import os

import numpy as np
import pytest

from oracle import collision_oracle as C
from oracle import mpc_oracle as O


@pytest.fixture(scope="module")
def fn(golden_dir):
    return np.load(os.path.join(golden_dir, "functions.npz"))


@pytest.fixture(scope="module")
def courses(golden_dir):
    return np.load(os.path.join(golden_dir, "courses.npz"))


def _smoothed(course):
    c = course.copy()
    O.smooth_yaw(c[:, 2])
    return c


def test_smooth_yaw(fn):
    for name in ["intersection", "roundabout", "multilane", "rand"]:
        got = O.smooth_yaw(fn[f"smooth_{name}_in"].copy())
        assert np.array_equal(got, fn[f"smooth_{name}_out"])


def test_index_xref_rollout_linearise(fn, courses):
    c = _smoothed(courses["intersection"])
    p = O.Params(dl=float(fn["step_dl"]))
    for k in range(len(fn["step_ok"])):
        n = int(fn["step_ncourse"][k])
        cx, cy, cyaw = c[:n, 0], c[:n, 1], c[:n, 2]
        x, y, v, yaw = fn["step_state"][k]
        start = int(fn["step_start"][k])
        assert O.nearest_index_forward(x, y, cx, cy, start) == fn["step_near"][k]
        xref, target, reach = O.ref_trajectory(p, x, y, v, cx, cy, cyaw, start)
        assert target == fn["step_target"][k]
        assert np.array_equal(reach, fn["step_reach"][k])
        assert np.array_equal(xref, fn["step_xref"][k])
        xbar = O.rollout(p, fn["step_state"][k], fn["step_oa"][k], fn["step_od"][k])
        np.testing.assert_allclose(xbar, fn["step_xbar"][k], rtol=0, atol=1e-12)
        for t in range(p.T):
            A, B, Cc = O.linear_model(p, xbar[2, t], xbar[3, t])
            np.testing.assert_allclose(A, fn["step_A"][k][t], rtol=0, atol=1e-12)
            np.testing.assert_allclose(B, fn["step_B"][k][t], rtol=0, atol=1e-12)
            np.testing.assert_allclose(Cc, fn["step_C"][k][t], rtol=0, atol=1e-12)


def test_projector(fn):
    for a, ref in zip(fn["proj_angle"], fn["proj_out"]):
        np.testing.assert_allclose(O.projector(a), ref, rtol=0, atol=1e-15)


@pytest.mark.parametrize("name", ["intersection", "roundabout"])
def test_collision_flags(fn, courses, name):
    path = _smoothed(courses[name])
    geo = C.CarGeometry()
    dl = float(np.linalg.norm(path[0, :2] - path[1, :2]))
    margin = C.cutoff_margin(geo, dl)
    assert margin == int(fn[f"coll_{name}_margin"])
    fw = int(fn[f"coll_{name}_fw"])
    flags = fn[f"coll_{name}_flag"]
    assert 5 < flags.sum() < len(flags) - 5          # the fixture exercises both outcomes
    for k in range(len(flags)):
        idx, v = int(fn[f"coll_{name}_idx"][k]), float(fn[f"coll_{name}_v"][k])
        obs = fn[f"coll_{name}_obs"][k]
        pred = C.predict_obstacle(*obs[0], dt=0.2, L=geo.L)
        np.testing.assert_allclose(pred, fn[f"coll_{name}_pred0"][k], rtol=0, atol=1e-12)
        ego = C.ego_prediction(path[idx:], v, 0.2, 2.0, 30 / 3.6)
        assert len(ego) == fn[f"coll_{name}_nres"][k]
        flag, cut = C.collision_cut(geo, path, idx, v, obs, dt=0.2, frame_window=fw, max_accel=2.0,
                                    max_speed=30 / 3.6, margin=margin)
        assert int(flag) == flags[k]
        assert cut == fn[f"coll_{name}_cut"][k]


@pytest.mark.parametrize("name,steps,nflags", [("intersection", 91, 35), ("roundabout", 116, 56)])
def test_episode_replay(golden_dir, name, steps, nflags):
    """Every recorded step of the reference's closed loop, replayed through the oracle step function."""
    e = np.load(os.path.join(golden_dir, f"episode_{name}.npz"))
    assert len(e["state"]) == steps and int(e["flag"].sum()) == nflags
    course = e["course_smoothed"]
    geo = C.CarGeometry()
    p = O.Params(dl=float(e["dl"]))
    for k in range(steps):
        n = int(e["ncourse"][k])
        flag, cut = C.collision_cut(geo, course, int(e["agent_idx"][k]), float(e["state"][k][2]), e["obs"][k],
                                    dt=0.2, frame_window=int(e["frame_window"]), max_accel=2.0, max_speed=30 / 3.6,
                                    margin=int(e["margin"]))
        assert int(flag) == e["flag"][k] and cut == n
        warm = bool(e["warm"][k])
        r = O.mpc_step(p, e["state"][k], e["oa_in"][k] if warm else None, e["od_in"][k] if warm else None,
                       course[:n, 0], course[:n, 1], course[:n, 2], int(e["target_in"][k]))
        assert r.status == O.STATUS_OPTIMAL
        assert r.target_ind == e["target_out"][k]
        assert np.array_equal(r.reaches_end, e["reach"][k])
        assert np.array_equal(r.xref, e["xref"][k])
        np.testing.assert_allclose(r.xbar, e["xbar"][k], rtol=0, atol=1e-12)
        for key, got in [("oa", r.oa), ("od", r.od), ("ox", r.ox), ("oy", r.oy), ("ov", r.ov), ("oyaw", r.oyaw)]:
            np.testing.assert_allclose(got, e[key][k], rtol=0, atol=1e-9)
        assert abs(r.cost - e["cost"][k]) <= 1e-9 * max(1.0, abs(e["cost"][k]))


def test_first_step_known_answer(golden_dir):
    """intersection(1,1), v0=0, zero warm start."""
    e = np.load(os.path.join(golden_dir, "episode_intersection.npz"))
    assert e["target_out"][0] == 1
    np.testing.assert_allclose(e["oa"][0][:12], 2.0, atol=1e-8)
    assert abs(e["oa"][0][12] - 1.0) < 0.05
    np.testing.assert_allclose(e["od"][0], 0.0, atol=1e-8)
    np.testing.assert_allclose(e["ov"][0][:13], 0.4 * np.arange(13), atol=1e-8)
    assert abs(e["cost"][0] - 71.9919) < 1e-3
