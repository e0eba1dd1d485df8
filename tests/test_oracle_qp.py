"""Row 8 (the QP, main/lib/mpc.py:141-211): the oracle's certified solve vs an independent solver.

cvxpy+ECOS are not installable offline, so the second opinion is the HiGHS convex-QP solver bundled with
scipy (private API, probed; the test is skipped when scipy moves it)."""
import os

import numpy as np
import pytest

from oracle import mpc_oracle as O
from oracle.qp import kkt_residuals, solve_qp


def _highs():
    try:
        from scipy.optimize._highspy import _core as hc
        hc.HighsHessian
        return hc
    except Exception:
        return None


def _solve_highs(hc, qp):
    import scipy.sparse as sp
    n = qp.P.shape[0]
    rows = np.vstack([qp.A, qp.G])
    lo = np.concatenate([qp.b, np.full(len(qp.h), -hc.kHighsInf)])
    hi = np.concatenate([qp.b, qp.h])
    lp = hc.HighsLp()
    lp.num_col_, lp.num_row_ = n, rows.shape[0]
    lp.col_cost_ = qp.q
    lp.col_lower_ = np.full(n, -hc.kHighsInf)
    lp.col_upper_ = np.full(n, hc.kHighsInf)
    lp.row_lower_, lp.row_upper_ = lo, hi
    lp.offset_ = qp.c0
    a = sp.csc_matrix(rows)
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.start_, lp.a_matrix_.index_, lp.a_matrix_.value_ = a.indptr, a.indices, a.data
    hs = hc.HighsHessian()
    hs.dim_ = n
    hs.format_ = hc.HessianFormat.kTriangular
    low = sp.csc_matrix(np.tril(qp.P))
    hs.start_, hs.index_, hs.value_ = low.indptr, low.indices, low.data
    model = hc.HighsModel()
    model.lp_, model.hessian_ = lp, hs
    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    h.passModel(model)
    h.run()
    assert h.getModelStatus() == hc.HighsModelStatus.kOptimal
    return np.array(h.getSolution().col_value), h.getInfo().objective_function_value


def test_qp_solver_small_known_answer():
    # min (x-2)^2 + (y-1)^2  s.t. x + y = 2, x <= 1.2  -> x = 1.2, y = 0.8
    P = 2 * np.eye(2)
    q = np.array([-4.0, -2.0])
    r = solve_qp(P, q, np.array([[1.0, 1.0]]), np.array([2.0]), np.array([[1.0, 0.0]]), np.array([1.2]), c0=5.0)
    assert r.ok
    np.testing.assert_allclose(r.z, [1.2, 0.8], atol=1e-10)
    assert abs(r.obj - (0.64 + 0.04)) < 1e-10


def test_oracle_qp_matches_highs(golden_dir):
    hc = _highs()
    if hc is None:
        pytest.skip("scipy-bundled HiGHS QP interface not available")
    e = np.load(os.path.join(golden_dir, "episode_intersection.npz"))
    p = O.Params(dl=float(e["dl"]))
    worst = 0.0
    for k in range(0, len(e["state"]), 6):
        qp = O.build_qp(p, e["xref"][k], e["xbar"][k], e["state"][k], e["reach"][k])
        res = solve_qp(qp.P, qp.q, qp.A, qp.b, qp.G, qp.h, c0=qp.c0)
        assert res.ok, res.kkt
        zh, fh = _solve_highs(hc, qp)
        worst = max(worst, np.abs(zh - res.z).max())
        # HiGHS' own tolerances are ~1e-7; the oracle is certified to 1e-9
        np.testing.assert_allclose(res.z, zh, rtol=1e-5, atol=1e-5)
        assert abs(res.obj - fh) <= 1e-6 * max(1.0, abs(fh))
        # the oracle point must be at least as good a KKT point as the cross-check's
        assert max(res.kkt.values()) <= 1e-9
    assert worst < 1e-5


def test_infeasible_predicate():
    p = O.Params()
    assert O.is_feasible(p, 30 / 3.6) and O.is_feasible(p, -5.0) and O.is_feasible(p, 0.0)
    assert not O.is_feasible(p, np.nextafter(30 / 3.6, 100.0)) and not O.is_feasible(p, -5.0000001)


def _workload_qps(w, idx):
    """(Params, SparseQP, certified oracle result) of workload instances: the QP the step builds after index / xref /
    rollout, exactly as the GPU parity tests feed it."""
    from helpers import default_vector, params_from_vector
    base = default_vector(w)
    for k in idx:
        pv = base if w.get("params") is None else w["params"][k]
        p = params_from_vector(pv, w["T"])
        cid = int(w["course_id"][k]) if w.get("course_id") is not None else 0
        c = w["courses"][cid][:int(w["course_len"][k])]
        x0 = w["state"][k]
        if not O.is_feasible(p, float(x0[2])):
            continue
        try:
            xref, _, reach = O.ref_trajectory(p, x0[0], x0[1], x0[2], c[:, 0], c[:, 1], c[:, 2], int(w["target_ind"][k]))
        except O.IndexRuleError:
            continue
        xbar = O.rollout(p, x0, w["oa"][k], w["od"][k])
        yield p, O.build_qp(p, xref, xbar, x0, reach)


@pytest.mark.parametrize("name", ["config2_T20", "config3_T13", "sweep_T8", "sweep_T25", "degenerate_T13"])
def test_oracle_qp_matches_highs_on_the_bench_workloads(name):
    """Row 8 beyond the one recorded episode: the certified oracle solve against HiGHS on instances of every
    workload family the GPU parity tests and bench.py use (warm starts, cut courses, per-instance parameters with both
    dt values, zero weights).  On the degenerate points the minimiser need not be unique: objective only."""
    hc = _highs()
    if hc is None:
        pytest.skip("scipy-bundled HiGHS QP interface not available")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from junction_mpc import synth
    w = {"config2_T20": lambda: synth.make_workload(2, B=256), "config3_T13": lambda: synth.make_workload(3, B=256),
         "sweep_T8": lambda: synth.make_sweep_sample(8, 256), "sweep_T25": lambda: synth.make_sweep_sample(25, 256),
         "degenerate_T13": lambda: synth.make_degenerate(13, B=256)}[name]()
    unique = not name.startswith("degenerate")
    n, worst_z, worst_f = 0, 0.0, 0.0
    for p, qp in _workload_qps(w, range(0, w["B"], 8)):
        res = solve_qp(qp.P, qp.q, qp.A, qp.b, qp.G, qp.h, c0=qp.c0)
        assert res.ok and max(res.kkt.values()) <= 1e-9, res.kkt
        zh, fh = _solve_highs(hc, qp)
        worst_f = max(worst_f, abs(res.obj - fh) / max(1.0, abs(fh)))
        if unique:
            # the controls (last 2T variables) and states, at the cross-check's own accuracy (~1e-7 residuals on
            # problems whose cost is nearly flat in some directions: DESIGN.md section 3)
            worst_z = max(worst_z, float(np.abs(zh - res.z).max()))
        # the certified point may not be worse than the cross-check's
        assert res.obj <= fh + 1e-6 * max(1.0, abs(fh))
        n += 1
    assert n >= 24
    assert worst_f <= 1e-6
    if unique:
        assert worst_z <= 1e-4, worst_z
