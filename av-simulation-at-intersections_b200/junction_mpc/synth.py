"""Seeded synthetic workloads for the batched MPC step.

Host-side numpy only; the same frozen inputs feed the CUDA path, the CPU oracle and the reference arm.
A workload is a dict of numpy arrays in the layout the C ABI takes (include/jmpc.h):

    state      [B, 4] float64   columns x, y, v, yaw
    course_id  [B]    int32     index into the course table
    course_len [B]    int32     effective course length N' (prefix truncation, mpc_intersection.py:138)
    target_ind [B]    int32     search start for the nearest-index rule
    oa, od     [B, T] float64   previous solution = linearisation point (zeros when there is none)
    params     [B, NPARAM] float64 or None   per-instance parameter overrides (config 5)
    obstacles  [B, n_obs, 6] float64 or None (x, y, v, yaw, a, steer) for the flag kernel
    agent_idx  [B] int32        ego index on the full course for the flag kernel

The index-rule validity filter of the generator (instances for which trajectories.py:120 would raise are
redrawn) is a pure numpy restatement of the 3-nearest rule and is part of input generation, not of the
product path.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import numpy as np

from .config import MPCConfig, PARAM_INDEX, NPARAM

_COURSE_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "courses.npz")


def smooth_yaw_inplace(yaw: np.ndarray) -> np.ndarray:
    """Yaw unwrapping of the course (reference main/lib/mpc.py:46-58), done once per course on the host."""
    for k in range(1, len(yaw)):
        while yaw[k] - yaw[k - 1] >= math.pi / 2.0:
            yaw[k] -= 2.0 * math.pi
        while yaw[k] - yaw[k - 1] <= -math.pi / 2.0:
            yaw[k] += 2.0 * math.pi
    return yaw


def load_course(name: str) -> np.ndarray:
    """Planner output (N x 3 float64: x, y, yaw) of the reference's A* search for the three scenario families,
    recorded by tests/golden/make_golden.py; yaw is returned smoothed, as MPC.__init__ leaves it."""
    with np.load(_COURSE_FILE) as z:
        c = np.array(z[name], dtype=np.float64)
    smooth_yaw_inplace(c[:, 2])
    return c


def _index_rule_ok(x, y, start, n, cx, cy, chunk: int = 8192) -> np.ndarray:
    """Vectorised 3-nearest rule of trajectories.py:100-126 on the window [start, n) of the course: True where the
    reference would return an index, False where it would raise."""
    ok = np.ones(len(x), bool)
    j = np.arange(len(cx))[None, :]
    for lo in range(0, len(x), chunk):
        sl = slice(lo, lo + chunk)
        d = np.hypot(cx[None, :] - x[sl, None], cy[None, :] - y[sl, None])
        d[(j < start[sl, None]) | (j >= n[sl, None])] = np.inf
        idx = np.argpartition(d, 2, axis=1)[:, :3]
        dd = np.take_along_axis(d, idx, axis=1)
        order = np.lexsort((idx, dd), axis=1)                      # by distance, ties by index
        idx = np.take_along_axis(idx, order, axis=1)
        good = (np.abs(idx[:, 1] - idx[:, 2]) == 2) | (np.abs(idx[:, 0] - idx[:, 1]) == 1)
        ok[sl] = good | ((n[sl] - start[sl]) < 3)
    return ok


def make_states(rng, course: np.ndarray, B: int, T: int, cut_fraction: float = 0.3) -> Dict[str, np.ndarray]:
    N = len(course)
    parts = []
    have = 0
    while have < B:
        m = int((B - have) * 1.1) + 16
        s = rng.integers(0, N - 60, m)
        x = course[s, 0] + rng.uniform(-0.5, 0.5, m)
        y = course[s, 1] + rng.uniform(-0.5, 0.5, m)
        yaw = course[s, 2] + rng.uniform(-0.1, 0.1, m)
        v = rng.uniform(0.0, 30.0 / 3.6, m)
        n_cut = rng.integers(s + 2, np.minimum(N, s + 300) + 1)
        n = np.where(rng.random(m) < cut_fraction, n_cut, N)
        t0 = np.maximum(s - 3, 0)
        keep = _index_rule_ok(x, y, t0, n, course[:, 0], course[:, 1])
        parts.append((np.stack([x, y, v, yaw], 1)[keep], t0[keep], n[keep], s[keep]))
        have += int(keep.sum())
    state = np.concatenate([p[0] for p in parts])[:B]
    target = np.concatenate([p[1] for p in parts])[:B].astype(np.int32)
    clen = np.concatenate([p[2] for p in parts])[:B].astype(np.int32)
    agent = np.concatenate([p[3] for p in parts])[:B].astype(np.int32)
    cold = rng.random(B) < 0.25
    oa = rng.uniform(-1.0, 2.0, (B, T))
    od = np.clip(np.cumsum(rng.uniform(-0.05, 0.05, (B, T)), axis=1) + rng.uniform(-0.2, 0.2, (B, 1)), -0.7, 0.7)
    oa[cold] = 0.0
    od[cold] = 0.0
    return dict(state=np.ascontiguousarray(state), target_ind=target, course_len=clen, agent_idx=agent, oa=oa, od=od,
                course_id=np.zeros(B, np.int32))


def make_obstacles(rng, B: int, n_obs: int) -> np.ndarray:
    obs = np.zeros((B, n_obs, 6))
    obs[:, :, 0] = rng.uniform(-35, 35, (B, n_obs))
    obs[:, :, 1] = rng.uniform(-35, 35, (B, n_obs))
    obs[:, :, 2] = rng.uniform(0, 30 / 3.6, (B, n_obs))
    obs[:, :, 3] = rng.uniform(-math.pi, math.pi, (B, n_obs))
    obs[:, :, 5] = rng.uniform(-0.4, 0.4, (B, n_obs))
    return obs


def make_workload(config: int, B: Optional[int] = None, cfg: Optional[MPCConfig] = None,
                  seed_offset: int = 0) -> Dict[str, object]:
    """Configs 2-5 of BASELINE.json (config 1 is the recorded golden episode under tests/golden/).
    `seed_offset` gives every rank of a multi-GPU run its own batch (seed = config when 0)."""
    cfg = cfg or MPCConfig.default()
    rng = np.random.default_rng(config if seed_offset == 0 else config * 1000 + seed_offset)
    if config == 2:
        B = B or 4096
        T = 20
        course = load_course("intersection")
        w = make_states(rng, course, B, T)
        w.update(T=T, courses=[course], params=None, obstacles=None, frame_window=10, name="intersection_T20")
    elif config in (3, 4):
        B = B or (65536 if config == 3 else 262144)
        T = 13
        course = load_course("roundabout" if config == 3 else "multilane")
        w = make_states(rng, course, B, T, cut_fraction=0.0)
        w.update(T=T, courses=[course], params=None, obstacles=make_obstacles(rng, B, 2 if config == 3 else 4),
                 frame_window=20, name="roundabout_T13" if config == 3 else "multilane_T13")
    elif config == 5:
        raise ValueError("config 5 is horizon-heterogeneous: use make_sweep()")
    else:
        raise ValueError(f"unknown config {config}")
    w["dl"] = float(np.linalg.norm(w["courses"][0][0, :2] - w["courses"][0][1, :2]))
    w["B"] = B
    return w


SWEEP_AXES = dict(
    T=[8, 13, 20, 25], dt=[0.1, 0.2], w_perp=[1., 10., 20., 50.], w_para=[0.1, 1., 5., 10.],
    R_acc=[.01, .1, 1., 10.], R_steer=[.01, .1, 1., 10.], Rd_acc=[1., 5., 10., 20.], Rd_steer=[.01, .1, 1., 10.])


def make_sweep(T: int, states_per_point: int = 32, max_points: Optional[int] = None,
               cfg: Optional[MPCConfig] = None, seed: int = 5) -> Dict[str, object]:
    """Config 5 restricted to one horizon (each launch is horizon-homogeneous): the full factorial grid of
    the reference's own sweep lists (mpc_sensitivity_analysis_comulative.py:103-128, zeros excluded) over
    dt x w_perp x w_para x R x Rd, `states_per_point` random states per grid point."""
    cfg = cfg or MPCConfig.default()
    rng = np.random.default_rng(seed * 1000 + T)
    axes = [SWEEP_AXES[k] for k in ["dt", "w_perp", "w_para", "R_acc", "R_steer", "Rd_acc", "Rd_steer"]]
    grid = np.array(np.meshgrid(*axes, indexing="ij")).reshape(len(axes), -1).T      # [8192, 7]
    if max_points is not None and max_points < grid.shape[0]:
        grid = grid[rng.choice(grid.shape[0], max_points, replace=False)]
    P = grid.shape[0]
    B = P * states_per_point
    course = load_course("intersection")
    w = make_states(rng, course, B, T)
    base = cfg.with_T(T).param_vector(dl=float(np.linalg.norm(course[0, :2] - course[1, :2])), dt=0.2,
                                      L=2.86, speed=30 / 3.6)
    params = np.repeat(base[None, :], B, axis=0)
    rep = np.repeat(grid, states_per_point, axis=0)
    for col, key in enumerate(["dt", "w_perp", "w_para", "R_a", "R_d", "Rd_a", "Rd_d"]):
        params[:, PARAM_INDEX[key]] = rep[:, col]
    w.update(T=T, courses=[course], params=params, obstacles=None, frame_window=10, name=f"sweep_T{T}", B=B,
             dl=float(base[PARAM_INDEX["dl"]]))
    return w



# ---- config 5 at full size: 4 horizons x 8192 parameter points x 32 states = 1 048 576 instances -----------------
SWEEP_HORIZONS = (8, 13, 20, 25)
SWEEP_CHUNKS = 16                       # a horizon slice is generated in 16 independent chunks of 16 384 instances


def sweep_grid() -> np.ndarray:
    """The 8192 parameter points of one horizon, [8192, 7] in the fixed order dt, w_perp, w_para, R_a, R_d, Rd_a, Rd_d."""
    axes = [SWEEP_AXES[k] for k in ["dt", "w_perp", "w_para", "R_acc", "R_steer", "Rd_acc", "Rd_steer"]]
    return np.array(np.meshgrid(*axes, indexing="ij")).reshape(len(axes), -1).T


def sweep_slice_size(states_per_point: int = 32) -> int:
    return sweep_grid().shape[0] * states_per_point


def make_sweep_shard(T: int, lo: int, hi: int, states_per_point: int = 32, cfg: Optional[MPCConfig] = None,
                     seed: int = 5) -> Dict[str, object]:
    """Instances [lo, hi) of the horizon-T slice of config 5 (instance i = parameter point i // states_per_point).
    The slice is defined chunk by chunk (each chunk has its own seeded generator), so a rank of a sharded run only
    generates the chunks its shard touches and every world size sees the same global batch."""
    cfg = cfg or MPCConfig.default()
    grid = sweep_grid()
    total = grid.shape[0] * states_per_point
    if not (0 <= lo <= hi <= total):
        raise ValueError("shard out of range")
    per_chunk = total // SWEEP_CHUNKS
    course = load_course("intersection")
    dl = float(np.linalg.norm(course[0, :2] - course[1, :2]))
    base = cfg.with_T(T).param_vector(dl=dl, dt=0.2, L=2.86, speed=30 / 3.6)
    parts = []
    for c in range(lo // per_chunk, (max(hi, lo + 1) - 1) // per_chunk + 1):
        rng = np.random.default_rng([seed, T, c])
        w = make_states(rng, course, per_chunk, T)
        first = c * per_chunk
        point = (first + np.arange(per_chunk)) // states_per_point
        params = np.repeat(base[None, :], per_chunk, axis=0)
        for col, key in enumerate(["dt", "w_perp", "w_para", "R_a", "R_d", "Rd_a", "Rd_d"]):
            params[:, PARAM_INDEX[key]] = grid[point, col]
        w["params"] = params
        a, b = max(lo, first) - first, min(hi, first + per_chunk) - first
        parts.append({k: v[a:b] for k, v in w.items()})
    out = {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in parts[0]}
    out.update(T=T, courses=[course], obstacles=None, frame_window=10, name=f"sweep_T{T}[{lo}:{hi}]", B=hi - lo, dl=dl)
    return out


def make_sweep_sample(T: int, n: int, states_per_point: int = 32, seed: int = 5) -> Dict[str, object]:
    """A bounded sample of the horizon-T slice of config 5 for the CPU legs: the first `n` states of the slice's
    generator, each on one of `n` parameter points strided evenly over the 8192-point grid."""
    w = make_sweep_shard(T, 0, n, states_per_point=states_per_point, seed=seed)
    grid = sweep_grid()
    pts = (np.arange(n) * (grid.shape[0] // max(n, 1))) % grid.shape[0]
    for col, key in enumerate(["dt", "w_perp", "w_para", "R_a", "R_d", "Rd_a", "Rd_d"]):
        w["params"][:, PARAM_INDEX[key]] = grid[pts, col]
    w["name"] = f"sweep_T{T}_sample{n}"
    return w


# The zero-valued points of the reference's sweep lists (mpc_sensitivity_analysis_comulative.py:103-128), each on the
# sensitivity base configuration (mpc_config_sensitivity.json: R = [0.1, 0.01], Rd = [10, 10]): a vanishing weight
# can leave the optimiser non-unique in the controls, so this side set is judged on cost and states only.
DEGENERATE_POINTS = [
    dict(w_perp=0.0), dict(w_para=0.0),
    dict(R_a=0.0, R_d=0.01), dict(R_a=0.1, R_d=0.0),
    dict(Rd_a=0.0, Rd_d=1.0), dict(Rd_a=10.0, Rd_d=0.0),
]


def make_degenerate(T: int, B: int = 1024, seed: int = 55) -> Dict[str, object]:
    """Config 5's degenerate side set, one horizon: B states spread over DEGENERATE_POINTS."""
    rng = np.random.default_rng(seed * 1000 + T)
    course = load_course("intersection")
    dl = float(np.linalg.norm(course[0, :2] - course[1, :2]))
    w = make_states(rng, course, B, T)
    sens = MPCConfig.from_dict({
        "NX": 4, "NU": 2, "T": T, "w_perp": 20.0, "w_para": 1.0, "R": [0.1, 0.01], "Rd": [10, 10], "Q_v_yaw": [0.0, 0.5],
        "Qf": [1.0, 1.0, 0.0, 0.5], "GOAL_DIS": 1.5, "STOP_SPEED": 0.1389, "MAX_TIME": 13.0, "MAX_ITER": 1, "DU_TH": 0.1,
        "MAX_DSTEER": 30.0, "MAX_ACCEL": 2.0, "MAX_DECEL": -10})
    base = sens.param_vector(dl=dl, dt=0.2, L=2.86, speed=30 / 3.6)
    params = np.repeat(base[None, :], B, axis=0)
    point = np.arange(B) % len(DEGENERATE_POINTS)
    for k, pt in enumerate(DEGENERATE_POINTS):
        for key, val in pt.items():
            params[point == k, PARAM_INDEX[key]] = val
    w.update(T=T, courses=[course], params=params, obstacles=None, frame_window=10, name=f"degenerate_T{T}", B=B, dl=dl,
             point=point)
    return w
