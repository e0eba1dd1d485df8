"""The sensitivity study as one batched run against History tables recorded from the
reference's own sweep loop (`mpc_sensitivity_analysis_comulative.py`, `lib.mpc_sensitivity.MPC`)."""
import csv
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sweep_reproduces_reference_histories(golden_dir, tmp_path):
    import __graft_entry__ as g
    g.build()
    from junction_mpc import synth
    from junction_mpc.sweep import REFERENCE_STUDIES, run_sensitivity_studies, save_studies, HISTORY_COLUMNS
    gold = np.load(os.path.join(golden_dir, "sensitivity_runs.npz"))
    course = synth.load_course("intersection")
    runs = run_sensitivity_studies(course)
    assert len(runs) == sum(len(v[1]) for v in REFERENCE_STUDIES.values()) == 30
    by = {(r.study, r.value): r for r in runs}
    # the sensitivity-default episode takes 74 steps (15 s of simulated time)
    for key, tag in [(("w_perp", 20.0), "default"), (("w_para", 1.0), "default"), (("Rd_acc", 10.0), "default"),
                     (("w_para", 0.1), "w_para_0p1"), (("R_acc", 10.0), "R_acc_10"), (("Rd_steer", 0.0), "Rd_steer_0")]:
        r, ref = by[key], gold[tag]
        assert r.goal_reached and r.history.shape == ref.shape, (key, r.history.shape, ref.shape)
        if tag == "Rd_steer_0":
            # degenerate point (steering unobservable while v = 0): states, not steer
            np.testing.assert_allclose(r.history[:, [0, 1, 2, 3, 4, 6]], ref[:, [0, 1, 2, 3, 4, 6]], rtol=0, atol=2e-4)
        else:
            np.testing.assert_allclose(r.history, ref, rtol=0, atol=1e-5)
    assert by[("w_perp", 20.0)].steps == 74
    files = save_studies(runs, str(tmp_path))
    assert len(files) == 12
    z = np.load(os.path.join(str(tmp_path), "w_para_histories.npz"))
    assert list(z["columns"]) == list(HISTORY_COLUMNS) and np.array_equal(z["values"], [0, 0.1, 1, 5, 10])
    np.testing.assert_allclose(z["value_1"], gold["w_para_0p1"], rtol=0, atol=1e-5)
    with open(os.path.join(str(tmp_path), "R_acc_summary.csv")) as f:
        rows = list(csv.DictReader(f))
    assert [float(r["value"]) for r in rows] == [0, 0.01, 0.1, 1, 10] and int(rows[-1]["steps"]) == 104
