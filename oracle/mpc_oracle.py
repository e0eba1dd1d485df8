"""Float64 numpy restatement of the reference MPC step (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Every function cites the reference lines (relative to /root/reference/) it follows.  State order
inside the MPC is [x, y, v, yaw] (main/lib/mpc.py:291).  The QP is kept in the reference's own sparse
(x, u) form -- 4(T+1) + 2T variables with explicit dynamics equalities -- so that the condensed
formulation used on the GPU is checked against an independent statement of the problem.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field, replace
from typing import Optional, Sequence

import numpy as np

from .qp import QPResult, solve_qp

# status words shared with include/jmpc.h
STATUS_OPTIMAL = 0
STATUS_MAX_ITER = 1
STATUS_INFEASIBLE = 2
STATUS_INDEX_RULE = 3


@dataclass(frozen=True)
class Params:
    """Everything a solve depends on besides state, warm start and course.

    Derivations follow main/lib/mpc.py:14-39 (Qf is scaled by T at :28, MAX_DSTEER converted to rad/s at
    :37) and the class constants of main/lib/simulation.py:23-25."""
    T: int = 13
    dt: float = 0.2
    dl: float = 0.083
    L: float = 2.86                                  # main/lib/car_dimensions.py:84
    speed: float = 30.0 / 3.6                        # QP speed cap (mpc.py:190); ctor default mpc.py:247
    w_perp: float = 20.0
    w_para: float = 1.0
    R: tuple = (0.01, 0.01)
    Rd: tuple = (0.01, 1.0)
    Q_v_yaw: tuple = (0.0, 0.5)
    Qf: tuple = (13.0, 13.0, 0.0, 6.5)               # already multiplied by T
    R_end: tuple = (10.0, 10.0)                      # mpc.py:181
    max_dsteer: float = math.radians(30.0)           # rad/s
    max_accel: float = 2.0
    max_decel: float = -10.0
    max_steer: float = float(np.deg2rad(45.0))       # simulation.py:23
    sim_max_speed: float = 30.0 / 3.6                # simulation.py:24 (rollout clamp, NOT `speed`)
    min_speed: float = -5.0                          # simulation.py:25
    v_ref_min: float = 10.0 / 3.6                    # mpc.py:99
    v_ref: float = 0.0                               # speed profile of mpc_with_speed.py:104,280-282: xref[2] = v_ref
    v_ref_cut: float = 1e9                           #   for course indices < v_ref_cut, 0 beyond (lib.mpc: always 0)
    goal_dis: float = 1.5
    stop_speed: float = 0.1389
    max_iter: int = 1

    @staticmethod
    def from_config(cfg: dict, **overrides) -> "Params":
        T = int(overrides.pop("T", cfg["T"]))
        base = dict(
            T=T,
            w_perp=float(cfg["w_perp"]), w_para=float(cfg["w_para"]),
            R=tuple(float(v) for v in cfg["R"]), Rd=tuple(float(v) for v in cfg["Rd"]),
            Q_v_yaw=tuple(float(v) for v in cfg["Q_v_yaw"]),
            Qf=tuple(float(v) * T for v in cfg["Qf"]),
            max_dsteer=float(np.deg2rad(cfg["MAX_DSTEER"])),
            max_accel=float(cfg["MAX_ACCEL"]), max_decel=float(cfg["MAX_DECEL"]),
            goal_dis=float(cfg["GOAL_DIS"]), stop_speed=float(cfg["STOP_SPEED"]),
            max_iter=int(cfg["MAX_ITER"]),
        )
        base.update(overrides)
        return Params(**base)

    @staticmethod
    def from_json(path: str, **overrides) -> "Params":
        with open(path, "r") as f:
            return Params.from_config(json.load(f), **overrides)


# ----------------------------------------------------------------------------------------------------
# row 2: yaw unwrapping                                                        main/lib/mpc.py:46-58
# ----------------------------------------------------------------------------------------------------
def smooth_yaw(yaw: np.ndarray) -> np.ndarray:
    """In-place unwrap so that consecutive differences fall in (-pi/2, pi/2)."""
    two_pi = math.pi * 2.0
    half_pi = math.pi / 2.0
    for k in range(1, len(yaw)):
        while yaw[k] - yaw[k - 1] >= half_pi:
            yaw[k] -= two_pi
        while yaw[k] - yaw[k - 1] <= -half_pi:
            yaw[k] += two_pi
    return yaw


# ----------------------------------------------------------------------------------------------------
# row 3: nearest forward index                                        main/lib/trajectories.py:100-126
# ----------------------------------------------------------------------------------------------------
class IndexRuleError(Exception):
    """The reference raises a bare Exception("something wrong") at trajectories.py:120."""


def nearest_index_forward(x: float, y: float, cx: np.ndarray, cy: np.ndarray, start: int) -> int:
    ex = cx[start:] - x
    ey = cy[start:] - y
    dist = np.sqrt(ex * ex + ey * ey)          # == np.linalg.norm([ex, ey], axis=0)
    m = len(dist)
    if m <= 1:
        return start
    if m == 2:
        return start + 1
    if m > 3:
        cand = np.argpartition(dist, 3)[:3]
        cand = cand[np.argsort(dist[cand])]
    else:
        cand = np.argsort(dist)
    i0, i1, i2 = (int(c) for c in cand)
    if abs(i1 - i2) == 2:
        return i0 + start
    if abs(i0 - i1) == 1:
        return max(i0, i1) + start
    raise IndexRuleError("nearest-index rule failed")


# ----------------------------------------------------------------------------------------------------
# row 4: reference sampling along the course                                  main/lib/mpc.py:89-112
# ----------------------------------------------------------------------------------------------------
def ref_trajectory(p: Params, x: float, y: float, v: float, cx, cy, cyaw, start: int,
                   ov: Optional[np.ndarray] = None, cv: Optional[np.ndarray] = None):
    n_course = len(cx)
    start = nearest_index_forward(x, y, cx, cy, start)
    if ov is None:
        ov = np.full(p.T + 1, max(v, p.v_ref_min))
    travel = np.cumsum(np.abs(ov) * p.dt)                 # sequential float64 adds, T+1 terms
    hop = np.rint(travel / p.dl).astype(np.int64)         # round-half-even
    idx = np.minimum(hop + start, n_course - 1)
    xref = np.zeros((4, p.T + 1))
    xref[0] = cx[idx]
    xref[1] = cy[idx]
    xref[3] = cyaw[idx]
    # 0 in lib.mpc (never tracked, mpc.py:107); cv[idx] in mpc_with_speed.py:104 -- either a speed array `cv` or the
    # two-level profile its set_trajectory_fromarray builds (v_ref before the cut index, 0 from there on, :280-282)
    xref[2] = np.where(idx < p.v_ref_cut, p.v_ref if cv is None else np.asarray(cv, float)[idx], 0.0)
    reaches_end = idx == n_course - 1
    return xref, int(start), reaches_end


# ----------------------------------------------------------------------------------------------------
# rows 5 / 12: plant step and operating-point rollout
#   main/lib/simulation.py:35-47, main/bicycle/main.py:28-41, main/lib/mpc.py:115-129
# ----------------------------------------------------------------------------------------------------
def plant_step(p: Params, st: Sequence[float], a: float, delta: float, dt: Optional[float] = None):
    """st = (x, y, v, yaw) -> next (x, y, v, yaw).  Pose moves with the OLD speed; speed is clamped."""
    dt = p.dt if dt is None else dt
    x, y, v, yaw = st
    delta = max(min(delta, p.max_steer), -p.max_steer)
    x_dot = v * np.cos(yaw)
    y_dot = v * np.sin(yaw)
    yaw_dot = (v / p.L) * np.tan(delta)
    x = x + x_dot * dt
    y = y + y_dot * dt
    yaw = yaw + yaw_dot * dt
    v = v + a * dt
    v = max(min(v, p.sim_max_speed), p.min_speed)
    return float(x), float(y), float(v), float(yaw)


def rollout(p: Params, x0: Sequence[float], oa: Sequence[float], od: Sequence[float]) -> np.ndarray:
    xbar = np.zeros((4, p.T + 1))
    st = tuple(float(c) for c in x0)
    xbar[:, 0] = st
    for t in range(p.T):
        st = plant_step(p, st, float(oa[t]), float(od[t]))
        xbar[:, t + 1] = st
    return xbar


# ----------------------------------------------------------------------------------------------------
# rows 6 / 7: Jacobians at (v, phi, delta=0) and the oriented xy weight
#   main/lib/mpc.py:61-82 and :132-138
# ----------------------------------------------------------------------------------------------------
def linear_model(p: Params, v: float, phi: float, delta: float = 0.0):
    dt, L = p.dt, p.L
    A = np.eye(4)
    A[0, 2] = dt * math.cos(phi)
    A[0, 3] = -dt * v * math.sin(phi)
    A[1, 2] = dt * math.sin(phi)
    A[1, 3] = dt * v * math.cos(phi)
    A[3, 2] = dt * math.tan(delta) / L
    B = np.zeros((4, 2))
    B[2, 0] = dt
    B[3, 1] = dt * v / (L * math.cos(delta) ** 2)
    C = np.zeros(4)
    C[0] = dt * v * math.sin(phi) * phi
    C[1] = -dt * v * math.cos(phi) * phi
    C[3] = -dt * v * delta / (L * math.cos(delta) ** 2)
    return A, B, C


def projector(angle: float) -> np.ndarray:
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c * c, c * s], [c * s, s * s]])


def stage_state_weight(p: Params, psi: float, at_end: bool) -> np.ndarray:
    """4x4 state weight of stage t>=1 (mpc.py:160-173)."""
    Q = np.zeros((4, 4))
    if at_end:
        Q[np.diag_indices(4)] = p.Qf
    else:
        Q[:2, :2] = projector(psi + 0.5 * np.pi) * p.w_perp + projector(psi) * p.w_para
        Q[2, 2], Q[3, 3] = p.Q_v_yaw
    return Q


# ----------------------------------------------------------------------------------------------------
# row 8: the QP in the reference's sparse (x, u) form                         main/lib/mpc.py:141-211
# ----------------------------------------------------------------------------------------------------
@dataclass
class SparseQP:
    P: np.ndarray
    q: np.ndarray
    c0: float
    A: np.ndarray
    b: np.ndarray
    G: np.ndarray
    h: np.ndarray
    T: int

    def ix(self, i: int, t: int) -> int:      # x[i, t]
        return 4 * t + i

    def iu(self, i: int, t: int) -> int:      # u[i, t]
        return 4 * (self.T + 1) + 2 * t + i


def is_feasible(p: Params, v0: float) -> bool:
    """The QP is feasible iff the (equality-fixed) initial speed satisfies its own bound rows
    (mpc.py:189-191 at t=0): take a == 0, delta == 0 for the rest."""
    return bool(p.min_speed <= v0 <= p.speed)


def build_qp(p: Params, xref: np.ndarray, xbar: np.ndarray, x0: Sequence[float], reaches_end) -> SparseQP:
    T = p.T
    nz = 4 * (T + 1) + 2 * T
    qp = SparseQP(P=np.zeros((nz, nz)), q=np.zeros(nz), c0=0.0, A=None, b=None, G=None, h=None, T=T)
    P, q = qp.P, qp.q
    # cost = sum_t (xref_t - x_t)' Q_t (xref_t - x_t) + u' R u + du' Rd du   (no 1/2 anywhere)
    for t in range(1, T + 1):
        Q = stage_state_weight(p, float(xref[3, t]), bool(reaches_end[t]))
        sl = slice(4 * t, 4 * t + 4)
        P[sl, sl] += 2.0 * Q
        q[sl] += -2.0 * Q @ xref[:, t]
        qp.c0 += float(xref[:, t] @ Q @ xref[:, t])
    for t in range(T):
        r = p.R_end if reaches_end[t] else p.R
        for i in range(2):
            P[qp.iu(i, t), qp.iu(i, t)] += 2.0 * r[i]
    for t in range(T - 1):
        for i in range(2):
            a, b_ = qp.iu(i, t), qp.iu(i, t + 1)
            P[a, a] += 2.0 * p.Rd[i]
            P[b_, b_] += 2.0 * p.Rd[i]
            P[a, b_] -= 2.0 * p.Rd[i]
            P[b_, a] -= 2.0 * p.Rd[i]
    # equalities: x_0 = x0 and x_{t+1} = A_t x_t + B_t u_t + C_t
    A = np.zeros((4 * (T + 1), nz))
    b = np.zeros(4 * (T + 1))
    A[:4, :4] = np.eye(4)
    b[:4] = x0
    for t in range(T):
        At, Bt, Ct = linear_model(p, float(xbar[2, t]), float(xbar[3, t]), 0.0)
        rows = slice(4 * (t + 1), 4 * (t + 1) + 4)
        A[rows, 4 * (t + 1):4 * (t + 1) + 4] = np.eye(4)
        A[rows, 4 * t:4 * t + 4] = -At
        A[rows, qp.iu(0, t):qp.iu(0, t) + 2] = -Bt
        b[rows] = Ct
    # inequalities G z <= h.  The t=0 speed rows act on an equality-fixed variable: they are the
    # feasibility predicate (is_feasible) and are left out of the solve.
    rows, rhs = [], []

    def add(coefs, bound):
        r = np.zeros(nz)
        for j, c in coefs:
            r[j] = c
        rows.append(r)
        rhs.append(bound)

    for t in range(1, T + 1):
        add([(qp.ix(2, t), 1.0)], p.speed)
        add([(qp.ix(2, t), -1.0)], -p.min_speed)
    for t in range(T):
        add([(qp.iu(0, t), 1.0)], p.max_accel)
        add([(qp.iu(0, t), -1.0)], -p.max_decel)
        add([(qp.iu(1, t), 1.0)], p.max_steer)
        add([(qp.iu(1, t), -1.0)], p.max_steer)
    lim = p.max_dsteer * p.dt
    for t in range(T - 1):
        add([(qp.iu(1, t + 1), 1.0), (qp.iu(1, t), -1.0)], lim)
        add([(qp.iu(1, t + 1), -1.0), (qp.iu(1, t), 1.0)], lim)
    qp.A, qp.b, qp.G, qp.h = A, b, np.array(rows), np.array(rhs)
    return qp


@dataclass
class StepResult:
    status: int
    target_ind: int
    xref: np.ndarray                      # (4, T+1)
    reaches_end: np.ndarray               # (T+1,) bool
    xbar: Optional[np.ndarray] = None     # (4, T+1)
    oa: Optional[np.ndarray] = None       # (T,)
    od: Optional[np.ndarray] = None
    ox: Optional[np.ndarray] = None       # (T+1,)
    oy: Optional[np.ndarray] = None
    ov: Optional[np.ndarray] = None
    oyaw: Optional[np.ndarray] = None
    cost: float = float("nan")
    qp: Optional[QPResult] = None


def linear_mpc_control(p: Params, xref, xbar, x0, reaches_end, tol: float = 1e-9):
    """Returns (status, oa, od, ox, oy, oyaw, ov, cost, QPResult)."""
    if not is_feasible(p, float(x0[2])):
        return STATUS_INFEASIBLE, None, None, None, None, None, None, float("nan"), None
    qp = build_qp(p, xref, xbar, x0, reaches_end)
    res = solve_qp(qp.P, qp.q, qp.A, qp.b, qp.G, qp.h, c0=qp.c0, tol=tol)
    T = p.T
    X = res.z[:4 * (T + 1)].reshape(T + 1, 4).T
    U = res.z[4 * (T + 1):].reshape(T, 2).T
    status = STATUS_OPTIMAL if res.ok else STATUS_MAX_ITER
    return status, U[0].copy(), U[1].copy(), X[0].copy(), X[1].copy(), X[3].copy(), X[2].copy(), res.obj, res


# ----------------------------------------------------------------------------------------------------
# row 9: one MPC step (MAX_ITER linearise->solve rounds)                      main/lib/mpc.py:214-242
# ----------------------------------------------------------------------------------------------------
def mpc_step(p: Params, x0: Sequence[float], oa, od, cx, cy, cyaw, target_ind: int, tol: float = 1e-9,
             cv: Optional[np.ndarray] = None, du_th: float = 0.0) -> StepResult:
    """x0 = (x, y, v, yaw).  oa/od = previous solution (unshifted) or None.  `cv`: reference speed per course point
    (mpc_with_speed.py:104).  `du_th` > 0 enables the exit the reference left commented out (mpc.py:236-240)."""
    if oa is None or od is None:
        oa = np.zeros(p.T)
        od = np.zeros(p.T)
    ov = None
    out = None
    for _ in range(p.max_iter):
        try:
            xref, target_ind, reaches_end = ref_trajectory(p, x0[0], x0[1], x0[2], cx, cy, cyaw, target_ind, ov, cv)
        except IndexRuleError:
            return StepResult(status=STATUS_INDEX_RULE, target_ind=int(target_ind), xref=np.zeros((4, p.T + 1)),
                              reaches_end=np.zeros(p.T + 1, bool))
        xbar = rollout(p, x0, oa, od)
        status, oa_n, od_n, ox, oy, oyaw, ov_n, cost, qres = linear_mpc_control(p, xref, xbar, x0, reaches_end, tol)
        out = StepResult(status=status, target_ind=target_ind, xref=xref, reaches_end=reaches_end, xbar=xbar,
                         oa=oa_n, od=od_n, ox=ox, oy=oy, ov=ov_n, oyaw=oyaw, cost=cost, qp=qres)
        if oa_n is None:
            # the reference would crash in the next round (np.abs(None)); MAX_ITER is 1 everywhere
            break
        du = float(np.sum(np.abs(oa_n - np.asarray(oa, float))) + np.sum(np.abs(od_n - np.asarray(od, float))))
        oa, od, ov = oa_n, od_n, ov_n
        if du_th > 0.0 and du <= du_th:               # mpc.py:236-240
            break
    return out


# ----------------------------------------------------------------------------------------------------
# row 10: the stateful controller                                            main/lib/mpc.py:245-330
# ----------------------------------------------------------------------------------------------------
class OracleMPC:
    """Same surface as the reference ``MPC`` class, built on the functions above."""

    def __init__(self, cx, cy, cyaw, dl, car_dimensions=None, speed: float = 30 / 3.6, dt: float = 0.2,
                 params: Optional[Params] = None):
        L = 2.86 if car_dimensions is None else float(car_dimensions.distance_back_to_front_wheel)
        self.params = replace(params or Params(), dl=float(dl), dt=float(dt), speed=float(speed), L=L)
        self.cx, self.cy = cx, cy
        self.cyaw = smooth_yaw(cyaw)
        self.dl, self.dt, self.speed = dl, dt, speed
        self.goal = (cx[-1], cy[-1])
        self.target_ind = 0
        self.oa = self.odelta = None
        self.ox = self.oy = self.oyaw = self.ov = self.xref = None
        self.di, self.ai = 0.0, 0.0
        self.last: Optional[StepResult] = None

    def set_trajectory_fromarray(self, trajectory: np.ndarray):
        self.cx, self.cy, self.cyaw = trajectory[:, 0], trajectory[:, 1], trajectory[:, 2]

    def step(self, state):
        x0 = [state.x, state.y, state.v, state.yaw]
        r = mpc_step(self.params, x0, self.oa, self.odelta, self.cx, self.cy, self.cyaw, self.target_ind)
        if r.status == STATUS_INDEX_RULE:
            raise IndexRuleError("something wrong")
        self.last = r
        self.oa, self.odelta, self.ox, self.oy, self.oyaw, self.ov = r.oa, r.od, r.ox, r.oy, r.oyaw, r.ov
        self.xref, self.target_ind = r.xref, r.target_ind
        if self.odelta is not None:
            self.di, self.ai = float(self.odelta[0]), float(self.oa[0])
        else:
            self.ai = self.params.max_decel
        return self.di, self.ai

    def get_current_xref_deviation(self) -> float:
        k = self.target_ind
        ex = self.cx[k] - self.ox[0]
        ey = self.cy[k] - self.oy[0]
        ang = self.cyaw[k] + np.pi / 2
        return float(np.linalg.norm(np.array([np.cos(ang) * ex, np.sin(ang) * ey])))

    def is_goal(self, state) -> bool:
        d = math.hypot(state.x - self.goal[0], state.y - self.goal[1])
        near = d <= self.params.goal_dis
        if abs(self.target_ind - len(self.cx)) >= 5:
            near = False
        return bool(near and abs(state.v) <= self.params.stop_speed)
