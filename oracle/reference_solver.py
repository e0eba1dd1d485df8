"""The reference's own solver as the arbiter of row 8 (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

``main/lib/mpc.py:141-211`` states the horizon QP with cvxpy and solves it with ECOS.  Both are third-party
(cvxpy >= 1.2.0, pyproject.toml:13; ecos >= 2.0.0, requirements.txt:9), not vendored, and absent from this image's
wheelhouse, so nothing here can run in the build container.  SURVEY.md section 8(c) asks for exactly this module:
probe ``import cvxpy, ecos`` where the tests and the bench run, and when both import, hand the frozen inputs to the
reference's formulation + ECOS and treat the answer as final.

* ``probe()``            -> {"available": bool, "reason": str, "versions": {...}}; never raises.
* ``solve_stage_qp(...)`` builds the cvxpy problem term by term as mpc.py:151-194 does (same variables x[4, T+1],
  u[2, T]; same quad_form costs; same constraint list) and calls ``prob.solve(solver=cvxpy.ECOS)`` as mpc.py:196-197.
  The model matrices come from ``oracle.mpc_oracle.linear_model`` (pinned to the reference's
  ``_get_linear_model_matrix`` by tests/golden/make_golden.py).
* ``mpc_step_reference_solver(...)`` is ``oracle.mpc_oracle.mpc_step`` with that solve in place of the certified
  one: what ``MPC.step`` of the reference computes (rows 3-7 of section 8(a) are pinned on the reference's own
  functions already).

``/root/reference`` does not exist on the GPU box, so the reference's mpc.py itself cannot be imported there; the
formulation above is its literal restatement on the cvxpy API, and the solver underneath is the real one.
"""
from __future__ import annotations

import importlib
from typing import Optional, Sequence

import numpy as np

from . import mpc_oracle as O

_PROBE = None


def probe() -> dict:
    """Is the reference's solver stack importable here?  Cached."""
    global _PROBE
    if _PROBE is not None:
        return _PROBE
    versions, missing = {}, []
    for name in ("cvxpy", "ecos"):
        try:
            m = importlib.import_module(name)
            versions[name] = getattr(m, "__version__", "unknown")
        except Exception as exc:                      # ImportError, or a broken install
            missing.append(f"{name}: {type(exc).__name__}")
    if missing:
        _PROBE = {"available": False, "versions": versions,
                  "reason": "import failed (" + "; ".join(missing) + "); not in the image, no network to install"}
    else:
        _PROBE = {"available": True, "versions": versions, "reason": "cvxpy + ecos import"}
    return _PROBE


def solve_stage_qp(p: O.Params, xref: np.ndarray, xbar: np.ndarray, x0: Sequence[float], reaches_end):
    """cvxpy + ECOS on the problem of mpc.py:141-211.  Returns (status_str, oa, od, ox, oy, oyaw, ov, objective)
    with None outputs when the status is neither OPTIMAL nor OPTIMAL_INACCURATE (mpc.py:199-209)."""
    import cvxpy
    T = p.T
    x = cvxpy.Variable((4, T + 1))
    u = cvxpy.Variable((2, T))
    R, Rd, R_end = np.diag(p.R), np.diag(p.Rd), np.diag(p.R_end)
    Q_v_yaw, Qf = np.diag(p.Q_v_yaw), np.diag(p.Qf)
    cost = 0.0
    cons = []
    for t in range(T + 1):
        if t > 0:                                                        # mpc.py:160-173
            if not reaches_end[t]:
                e_xy = xref[:2, t] - x[:2, t]
                cost += cvxpy.quad_form(e_xy, O.projector(xref[3, t] + 0.5 * np.pi) * p.w_perp)
                cost += cvxpy.quad_form(e_xy, O.projector(xref[3, t]) * p.w_para)
                cost += cvxpy.quad_form(xref[2:, t] - x[2:, t], Q_v_yaw)
            else:
                cost += cvxpy.quad_form(xref[:, t] - x[:, t], Qf)
        if t < T:                                                        # mpc.py:175-183
            A, B, C = O.linear_model(p, float(xbar[2, t]), float(xbar[3, t]), 0.0)
            cons.append(x[:, t + 1] == A @ x[:, t] + B @ u[:, t] + C)
            cost += cvxpy.quad_form(u[:, t], R_end if reaches_end[t] else R)
        if t < T - 1:                                                    # mpc.py:185-187
            cost += cvxpy.quad_form(u[:, t + 1] - u[:, t], Rd)
            cons.append(cvxpy.abs(u[1, t + 1] - u[1, t]) <= p.max_dsteer * p.dt)
    cons += [x[:, 0] == np.asarray(x0, float),                           # mpc.py:189-194
             x[2, :] <= p.speed, x[2, :] >= p.min_speed,
             u[0, :] <= p.max_accel, u[0, :] >= p.max_decel,
             cvxpy.abs(u[1, :]) <= p.max_steer]
    prob = cvxpy.Problem(cvxpy.Minimize(cost), cons)
    prob.solve(solver=cvxpy.ECOS, verbose=False)                         # mpc.py:196-197
    if prob.status in (cvxpy.OPTIMAL, cvxpy.OPTIMAL_INACCURATE):
        xv, uv = np.asarray(x.value), np.asarray(u.value)
        return (prob.status, uv[0].copy(), uv[1].copy(), xv[0].copy(), xv[1].copy(), xv[3].copy(), xv[2].copy(),
                float(prob.value))
    return prob.status, None, None, None, None, None, None, float("nan")


def mpc_step_reference_solver(p: O.Params, x0, oa, od, cx, cy, cyaw, target_ind: int,
                              cv: Optional[np.ndarray] = None) -> O.StepResult:
    """`MPC.step` of the reference with ITS solver: rows 3-7 from the (pinned) oracle functions, row 8 by cvxpy+ECOS."""
    if oa is None or od is None:
        oa, od = np.zeros(p.T), np.zeros(p.T)
    ov, out = None, None
    for _ in range(p.max_iter):
        try:
            xref, target_ind, reaches_end = O.ref_trajectory(p, x0[0], x0[1], x0[2], cx, cy, cyaw, target_ind, ov, cv)
        except O.IndexRuleError:
            return O.StepResult(status=O.STATUS_INDEX_RULE, target_ind=int(target_ind), xref=np.zeros((4, p.T + 1)),
                                reaches_end=np.zeros(p.T + 1, bool))
        xbar = O.rollout(p, x0, oa, od)
        st, oa_n, od_n, ox, oy, oyaw, ov_n, obj = solve_stage_qp(p, xref, xbar, x0, reaches_end)
        status = O.STATUS_OPTIMAL if oa_n is not None else O.STATUS_INFEASIBLE
        out = O.StepResult(status=status, target_ind=target_ind, xref=xref, reaches_end=reaches_end, xbar=xbar,
                           oa=oa_n, od=od_n, ox=ox, oy=oy, ov=ov_n, oyaw=oyaw, cost=obj)
        if oa_n is None:
            break
        oa, od, ov = oa_n, od_n, ov_n
    return out
