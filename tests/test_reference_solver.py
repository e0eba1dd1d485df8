"""Row 8 against the reference's own solver stack (cvxpy + ECOS), SURVEY.md section 8(c) last row.

The probe always runs; the comparisons run wherever both packages import (they are not in this image: the tests
then skip with the probe's reason, which bench.py also prints into its JSON line)."""
import os

import numpy as np
import pytest

from oracle import mpc_oracle as O
from oracle import reference_solver as RS


def test_probe_never_raises_and_reports():
    r = RS.probe()
    assert set(r) >= {"available", "reason", "versions"}
    assert isinstance(r["available"], bool) and r["reason"]
    print("reference solver probe:", r)


def _need_solver():
    r = RS.probe()
    if not r["available"]:
        pytest.skip("cvxpy + ECOS not importable here: " + r["reason"])


def test_oracle_matches_ecos_on_the_golden_episode(golden_dir):
    """The certified solve of oracle/qp.py vs ECOS on the recorded config-1 episode (every 3rd step)."""
    _need_solver()
    e = np.load(os.path.join(golden_dir, "episode_intersection.npz"))
    p = O.Params(dl=float(e["dl"]))
    for k in range(0, len(e["state"]), 3):
        st, oa, od, ox, oy, oyaw, ov, obj = RS.solve_stage_qp(p, e["xref"][k], e["xbar"][k], e["state"][k], e["reach"][k])
        assert oa is not None, st
        assert abs(obj - e["cost"][k]) <= 1e-4 * abs(e["cost"][k])
        np.testing.assert_allclose(oa, e["oa"][k], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(od, e["od"][k], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(ov, e["ov"][k], rtol=1e-3, atol=1e-4)


def test_first_step_known_answer_with_ecos(golden_dir):
    """intersection(1,1), v0 = 0, zero warm start -> oa = [2]*12 + [~1], objective 71.9919."""
    _need_solver()
    e = np.load(os.path.join(golden_dir, "episode_intersection.npz"))
    c = e["course_smoothed"]
    p = O.Params(dl=float(e["dl"]))
    r = RS.mpc_step_reference_solver(p, e["state"][0], None, None, c[:, 0], c[:, 1], c[:, 2], 0)
    assert r.target_ind == 1
    np.testing.assert_allclose(r.oa[:12], 2.0, atol=1e-5)
    assert abs(r.cost - 71.9919) < 1e-3
