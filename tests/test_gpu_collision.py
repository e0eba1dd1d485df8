"""Collision flag / cut index kernel against the fixtures recorded from the reference and against the oracle."""
import os

import numpy as np
import pytest

from oracle import collision_oracle as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jm():
    import __graft_entry__ as g
    g.build()
    from junction_mpc import synth
    from junction_mpc.batched import BatchedMPC
    return synth, BatchedMPC


@pytest.mark.parametrize("name", ["intersection", "roundabout"])
def test_flags_match_reference_fixture(jm, golden_dir, name):
    synth, BatchedMPC = jm
    fn = np.load(os.path.join(golden_dir, "functions.npz"))
    course = synth.load_course(name)
    mpc = BatchedMPC([course], dl=0.083, T=13, max_batch=256)
    flag, clen = mpc.collision_host(fn[f"coll_{name}_idx"], fn[f"coll_{name}_v"], fn[f"coll_{name}_obs"],
                                    frame_window=int(fn[f"coll_{name}_fw"]), margin=int(fn[f"coll_{name}_margin"]))
    assert np.array_equal(flag, fn[f"coll_{name}_flag"])
    assert np.array_equal(clen, fn[f"coll_{name}_cut"])


@pytest.mark.parametrize("name", ["intersection", "roundabout"])
def test_flags_match_reference_episode(jm, golden_dir, name):
    synth, BatchedMPC = jm
    e = np.load(os.path.join(golden_dir, f"episode_{name}.npz"))
    mpc = BatchedMPC([e["course_smoothed"]], dl=float(e["dl"]), T=13, max_batch=256)
    flag, clen = mpc.collision_host(e["agent_idx"], e["state"][:, 2], e["obs"], frame_window=int(e["frame_window"]),
                                    margin=int(e["margin"]))
    assert np.array_equal(flag, e["flag"])
    assert np.array_equal(clen, e["ncourse"])


@pytest.mark.parametrize("config,B", [(3, 2048), (4, 1024)])
def test_flags_match_oracle_on_synthetic(jm, config, B):
    synth, BatchedMPC = jm
    w = synth.make_workload(config, B=B)
    course = w["courses"][0]
    geo = C.CarGeometry()
    margin = C.cutoff_margin(geo, w["dl"])
    # pull half of the obstacles next to the path so that both outcomes occur
    rng = np.random.default_rng(99)
    near = rng.random(B) < 0.5
    k = np.minimum(w["agent_idx"] + rng.integers(0, 200, B), len(course) - 1)
    w["obstacles"][near, 0, 0] = course[k[near], 0] + rng.uniform(-8, 8, near.sum())
    w["obstacles"][near, 0, 1] = course[k[near], 1] + rng.uniform(-8, 8, near.sum())
    mpc = BatchedMPC([course], dl=w["dl"], T=13, max_batch=B)
    v = w["state"][:, 2].copy()
    v[::17] = 30 / 3.6                                   # exercises the constant-spacing branch
    flag, clen = mpc.collision_host(w["agent_idx"], v, w["obstacles"], frame_window=w["frame_window"], margin=margin)
    sel = np.arange(0, B, 8)
    for i in sel:
        f, n = C.collision_cut(geo, course, int(w["agent_idx"][i]), float(v[i]), w["obstacles"][i], dt=0.2,
                               frame_window=w["frame_window"], max_accel=2.0, max_speed=30 / 3.6, margin=margin)
        assert int(f) == flag[i] and n == clen[i], i
    assert 0.05 < flag.mean() < 0.95


def test_no_obstacles(jm):
    synth, BatchedMPC = jm
    course = synth.load_course("multilane")          # the reference's multi-lane script has moving_obstacles = []
    mpc = BatchedMPC([course], dl=0.083, T=13, max_batch=64)
    flag, clen = mpc.collision_host(np.arange(8, dtype=np.int32), np.ones(8), np.zeros((8, 0, 6)), 20, 72)
    assert (flag == 0).all() and (clen == len(course)).all()


def test_plant_step(jm):
    import torch
    synth, BatchedMPC = jm
    from helpers import params_from_vector
    from oracle import mpc_oracle as O
    w = synth.make_workload(2, B=64)
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=20, max_batch=64)
    rng = np.random.default_rng(0)
    a, d = rng.uniform(-10, 2, 64), rng.uniform(-1.0, 1.0, 64)
    st = torch.as_tensor(w["state"], device="cuda")
    mpc.plant_step(st, torch.as_tensor(a, device="cuda"), torch.as_tensor(d, device="cuda"))
    got = st.cpu().numpy()
    p = params_from_vector(mpc.default_params, 20)
    for k in range(64):
        np.testing.assert_allclose(got[k], O.plant_step(p, w["state"][k], a[k], d[k]), rtol=0, atol=1e-12)
