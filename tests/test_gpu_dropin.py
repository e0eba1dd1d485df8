"""The drop-in `MPC` class in closed loop: the reference's scenario loop (mpc_intersection.py:99-163) driven by
junction_mpc.mpc.MPC must retrace the recorded episode (91 steps, 35 collision flags for config 1)."""
import os
import sys
import types
from dataclasses import dataclass

import numpy as np
import pytest

from oracle import collision_oracle as C
from oracle import mpc_oracle as O

pytestmark = pytest.mark.gpu


@dataclass
class State:
    x: float = 0.0
    y: float = 0.0
    yaw: float = 0.0
    v: float = 0.0


class Car:
    distance_back_to_front_wheel = 2.86


@pytest.mark.parametrize("name", ["intersection", "roundabout"])
def test_closed_loop_retraces_reference_episode(golden_dir, name):
    import __graft_entry__ as g
    g.build()
    from junction_mpc.mpc import MPC
    e = np.load(os.path.join(golden_dir, f"episode_{name}.npz"))
    raw = np.load(os.path.join(golden_dir, "courses.npz"))[name]
    trajectory_full = raw.copy()
    dl = np.linalg.norm(trajectory_full[0, :2] - trajectory_full[1, :2])
    mpc = MPC(cx=trajectory_full[:, 0], cy=trajectory_full[:, 1], cyaw=trajectory_full[:, 2], dl=dl, dt=0.2,
              car_dimensions=Car(), speed=30 / 3.6)
    # the constructor smoothed the caller's yaw column in place (mpc.py:260)
    assert np.array_equal(trajectory_full, e["course_smoothed"])
    geo = C.CarGeometry()
    p = O.Params(dl=float(dl))
    margin, fw = int(e["margin"]), int(e["frame_window"])
    state = State(x=trajectory_full[0, 0], y=trajectory_full[0, 1], yaw=trajectory_full[0, 2], v=0.0)
    agent_idx, tmp, steps, flags = 0, None, 0, 0
    for i in range(400):
        if mpc.is_goal(state):
            break
        if tmp is None or np.any(tmp[agent_idx, :] != tmp[-1, :]):
            agent_idx = O.nearest_index_forward(state.x, state.y, trajectory_full[:, 0], trajectory_full[:, 1], agent_idx)
        # obstacles are replayed from the recording (their scripted motion is not part of the hot path)
        flag, cut = C.collision_cut(geo, trajectory_full, agent_idx, state.v, e["obs"][i], dt=0.2, frame_window=fw,
                                    max_accel=2.0, max_speed=30 / 3.6, margin=margin)
        assert int(flag) == e["flag"][i] and cut == e["ncourse"][i], i
        tmp = trajectory_full[:cut]
        mpc.set_trajectory_fromarray(tmp)
        np.testing.assert_allclose([state.x, state.y, state.v, state.yaw], e["state"][i], rtol=0, atol=1e-6)
        delta, acc = mpc.step(state)
        assert isinstance(delta, float) and isinstance(acc, float)
        assert mpc.xref.shape == (4, 14) and mpc.ox.shape == (14,)
        assert abs(delta - e["di"][i]) <= 1e-4 + 1e-3 * abs(e["di"][i])
        assert abs(acc - e["ai"][i]) <= 1e-4 + 1e-3 * abs(e["ai"][i])
        assert abs(mpc.get_current_xref_deviation() - e["dev"][i]) <= 1e-6
        flags += int(flag)
        steps += 1
        x, y, v, yaw = O.plant_step(p, (state.x, state.y, state.v, state.yaw), acc, delta)
        state = State(x=x, y=y, yaw=yaw, v=v)
    assert steps == len(e["state"]) and flags == int(e["flag"].sum())
    np.testing.assert_allclose([state.x, state.y, state.v, state.yaw], e["final_state"], rtol=0, atol=1e-5)


def test_install_registers_lib_mpc():
    from junction_mpc import mpc as M
    saved = {k: sys.modules.get(k) for k in ["lib", "lib.mpc", "lib.mpc_sensitivity"]}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        M.install()
        from lib.mpc import MPC, MAX_ACCEL          # the import line of mpc_intersection.py:20
        from lib.mpc_sensitivity import MPC as SMPC, MAX_ACCEL as A2
        assert MAX_ACCEL == 2.0 and A2 == 2.0 and MPC.__name__ == "MPC"
        assert "speed" not in SMPC.__init__.__code__.co_varnames[:7]
        import lib.mpc as lm
        assert lm.T == 13 and lm.Qf[0, 0] == 13.0 and lm.MAX_DECEL == -10
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_infeasible_single_instance_matches_reference_failure_path(capsys):
    from junction_mpc.mpc import MPC, MAX_DECEL
    from junction_mpc import synth
    c = synth.load_course("intersection")
    mpc = MPC(cx=c[:, 0], cy=c[:, 1], cyaw=c[:, 2], dl=0.083, car_dimensions=Car(), speed=5.0)
    di, ai = mpc.step(State(x=c[10, 0], y=c[10, 1], yaw=c[10, 2], v=6.0))      # v0 above the cap
    assert ai == MAX_DECEL and di == 0.0 and mpc.oa is None and mpc.ox is None
    assert "Cannot solve mpc" in capsys.readouterr().err
    di, ai = mpc.step(State(x=c[10, 0], y=c[10, 1], yaw=c[10, 2], v=4.0))      # recovers with a cold start
    assert mpc.oa is not None and len(mpc.oa) == 13


def test_speed_profile_flavour_retraces_reference_episode(golden_dir):
    """`lib.mpc_with_speed.MPC` in the loop of mpc_intersection_new_ref.py."""
    import __graft_entry__ as g
    g.build()
    from junction_mpc import mpc as M
    e = np.load(os.path.join(golden_dir, "episode_new_ref.npz"))
    raw = np.load(os.path.join(golden_dir, "courses.npz"))["intersection"]
    trajectory_full = raw.copy()
    dl = np.linalg.norm(trajectory_full[0, :2] - trajectory_full[1, :2])
    cv = np.full(trajectory_full[:, 1].shape, 30 / 3.6)
    mpc = M._WithSpeedMPC(cx=trajectory_full[:, 0], cy=trajectory_full[:, 1], cv=cv, cyaw=trajectory_full[:, 2], dl=dl,
                          dt=0.2, car_dimensions=Car())
    geo = C.CarGeometry()
    p = O.Params(dl=float(dl))
    margin, fw = int(e["margin"]), int(e["frame_window"])
    state = State(x=trajectory_full[0, 0], y=trajectory_full[0, 1], yaw=trajectory_full[0, 2], v=0.0)
    agent_idx, steps = 0, 0
    for i in range(400):
        if mpc.is_goal(state):
            break
        # the course is never truncated in this scenario, so the index update always runs
        agent_idx = O.nearest_index_forward(state.x, state.y, trajectory_full[:, 0], trajectory_full[:, 1], agent_idx)
        flag, cut = C.collision_cut(geo, trajectory_full, agent_idx, state.v, e["obs"][i], dt=0.2, frame_window=fw,
                                    max_accel=2.0, max_speed=30 / 3.6, margin=margin)
        cutoff = cut if flag else 999
        assert int(flag) == e["flag"][i] and cutoff == e["cutoff"][i], i
        mpc.set_trajectory_fromarray(trajectory_full, cutoff_idx=cutoff)
        np.testing.assert_allclose([state.x, state.y, state.v, state.yaw], e["state"][i], rtol=0, atol=1e-6)
        delta, acc = mpc.step(state)
        assert np.array_equal(mpc.xref, e["xref"][i])                 # includes the speed row cv[idx]
        assert abs(delta - e["di"][i]) <= 1e-4 + 1e-3 * abs(e["di"][i])
        assert abs(acc - e["ai"][i]) <= 1e-4 + 1e-3 * abs(e["ai"][i])
        assert abs(mpc.cost - e["cost"][i]) <= 1e-4 * abs(e["cost"][i])
        steps += 1
        x, y, v, yaw = O.plant_step(p, (state.x, state.y, state.v, state.yaw), acc, delta)
        state = State(x=x, y=y, yaw=yaw, v=v)
    assert steps == len(e["state"]) == 88
    np.testing.assert_allclose([state.x, state.y, state.v, state.yaw], e["final_state"], rtol=0, atol=1e-5)
    assert e["xref"][:, 2].max() == 25 / 3.6 and (e["xref"][:, 2] == 0).any()      # both levels of the profile occur
