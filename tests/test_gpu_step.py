"""Parity of the CUDA MPC step (through the C ABI) with the CPU oracle.  Needs a GPU: `pytest -m gpu`.

Gates (BASELINE.json north_star): controls and predicted states |d| <= 1e-4 + 1e-3 |ref|, cost 1e-4 relative,
target index / xref / status exact."""
import os

import numpy as np
import pytest

from helpers import compare_step, default_vector, oracle_batch, params_from_vector, scaled_err
from oracle import mpc_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jm():
    import __graft_entry__ as g
    g.build()
    from junction_mpc import synth
    from junction_mpc.batched import BatchedMPC
    return synth, BatchedMPC


def _run(BatchedMPC, w, idx=None, **kw):
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=max(w["B"], 64), **kw)
    sel = slice(None) if idx is None else idx
    out = mpc.step_host(w["state"][sel], w["target_ind"][sel], w["oa"][sel], w["od"][sel],
                        course_len=w["course_len"][sel], params=None if w["params"] is None else w["params"][sel])
    return mpc, out


def test_config2_subset_matches_oracle(jm):
    synth, BatchedMPC = jm
    w = synth.make_workload(2, B=4096)
    idx = np.arange(0, 4096, 8)                      # 512 instances, oracle in seconds
    mpc, out = _run(BatchedMPC, w)
    refs = oracle_batch(w, idx)
    worst = compare_step(out, refs, idx)
    assert worst <= 1.0
    assert (out.status == 0).all()
    assert out.iters.max() <= 40


@pytest.mark.parametrize("T", [8, 13, 20, 25])
def test_sweep_subset_matches_oracle(jm, T):
    """Config 5: per-instance parameters from the reference's sweep lists, all four horizons."""
    synth, BatchedMPC = jm
    w = synth.make_sweep(T, states_per_point=1, max_points=192)
    mpc, out = _run(BatchedMPC, w)
    refs = oracle_batch(w, range(w["B"]))
    assert compare_step(out, refs, range(w["B"])) <= 1.0


@pytest.mark.parametrize("name", ["intersection", "roundabout"])
def test_golden_episode(jm, golden_dir, name):
    """Config 1: every step the reference's closed loop took, solved as one batch."""
    synth, BatchedMPC = jm
    e = np.load(os.path.join(golden_dir, f"episode_{name}.npz"))
    course = e["course_smoothed"]
    B, T = e["state"].shape[0], e["oa"].shape[1]
    mpc = BatchedMPC([course], dl=float(e["dl"]), T=T, max_batch=256)
    out = mpc.step_host(e["state"], e["target_in"], e["oa_in"], e["od_in"], course_len=e["ncourse"], warm=e["warm"])
    assert (out.status == 0).all()
    assert np.array_equal(out.target_ind, e["target_out"])
    assert np.array_equal(out.xref, e["xref"])
    for key, got in [("oa", out.oa), ("od", out.od), ("ox", out.ox), ("oy", out.oy), ("ov", out.ov), ("oyaw", out.oyaw)]:
        assert scaled_err(got, e[key]) <= 1.0, key
    assert np.max(np.abs(out.cost - e["cost"]) / np.abs(e["cost"])) <= 1e-4
    # the controls the scenario loop actually applied
    assert scaled_err(out.od[:, 0], e["di"]) <= 1.0 and scaled_err(out.oa[:, 0], e["ai"]) <= 1.0


def test_edge_cases(jm):
    synth, BatchedMPC = jm
    w = synth.make_workload(2, B=16)
    c = w["courses"][0]
    N = len(c)
    st, tgt, clen = w["state"].copy(), w["target_ind"].copy(), w["course_len"].copy()
    # 0: infeasible (v0 above the cap); 1: v0 exactly at the cap (feasible, boundary); 2: v0 exactly MIN_SPEED
    st[0, 2] = 30 / 3.6 + 1e-9
    st[1, 2] = 30 / 3.6
    st[2, 2] = -5.0
    # 3..5: course truncated to 1, 2, 3 points past the search start (len <= 3 branches of the index rule)
    for k, extra in [(3, 1), (4, 2), (5, 3)]:
        clen[k] = tgt[k] + extra
    # 6: search start at the very end of the course
    tgt[6] = N - 1
    clen[6] = N
    st[6, :2] = c[N - 1, :2]
    # 7: far from the course start with the search window at 0 -> whatever the rule says, GPU == oracle
    st[7, :2] = c[300, :2] + 5.0
    tgt[7] = 0
    clen[7] = N
    w.update(state=st, target_ind=tgt, course_len=clen)
    mpc, out = _run(BatchedMPC, w)
    refs = oracle_batch(w, range(16), processes=1)
    assert refs[0].status == O.STATUS_INFEASIBLE and out.status[0] == O.STATUS_INFEASIBLE
    assert out.status[1] == O.STATUS_OPTIMAL and out.status[2] == O.STATUS_OPTIMAL
    # infeasible: control block untouched, target/xref still reported
    assert np.array_equal(out.oa[0], w["oa"][0]) and np.array_equal(out.od[0], w["od"][0])
    compare_step(out, refs, range(16))


def test_index_rule_failure_is_reported(jm):
    synth, BatchedMPC = jm
    # a course that doubles back: the three nearest points are not neighbours -> reference raises
    xs = np.concatenate([np.linspace(0, 10, 101), np.linspace(10, 0, 101)])
    ys = np.concatenate([np.zeros(101), np.full(101, 0.05)])
    course = np.stack([xs, ys, np.zeros(202)], axis=1)
    mpc = BatchedMPC([course], dl=0.1, T=13, max_batch=64)
    state = np.array([[5.0, 0.02, 1.0, 0.0], [5.0, 0.0, 1.0, 0.0]])
    out = mpc.step_host(state, np.zeros(2, np.int32))
    p = params_from_vector(mpc.default_params, 13)
    for k in range(2):
        r = O.mpc_step(p, state[k], None, None, course[:, 0], course[:, 1], course[:, 2], 0)
        assert int(out.status[k]) == r.status
    assert (out.status == O.STATUS_INDEX_RULE).any()
    k = int(np.nonzero(out.status == O.STATUS_INDEX_RULE)[0][0])
    assert out.target_ind[k] == 0                      # untouched


def test_iterative_linearisation(jm):
    """MAX_ITER = 3: the ov feedback into the reference sampling."""
    synth, BatchedMPC = jm
    w = synth.make_workload(2, B=48)
    mpc, out = _run(BatchedMPC, w, linearisation_iters=3)
    refs = oracle_batch(w, range(48), max_iter=3)
    assert compare_step(out, refs, range(48)) <= 1.0


def test_device_path_equals_host_path(jm):
    import torch
    synth, BatchedMPC = jm
    w = synth.make_workload(2, B=300)
    mpc, host = _run(BatchedMPC, w)
    dev = torch.device("cuda", 0)
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    state, tgt = t(w["state"], torch.float64), t(w["target_ind"], torch.int32)
    oa, od, clen = t(w["oa"], torch.float64), t(w["od"], torch.float64), t(w["course_len"], torch.int32)
    out = mpc.step(state, tgt, oa, od, mpc.alloc_outputs(300), course_len=clen)
    torch.cuda.synchronize()
    for key in ["oa", "od", "ox", "oy", "ov", "oyaw", "xref", "cost", "status", "target_ind"]:
        assert np.array_equal(getattr(out, key).cpu().numpy(), getattr(host, key)), key


def _check_properties(w, out, default_vector_fn):
    """Size-independent checks on a whole batch: every instance solved, all constraint rows satisfied, the
    linearised speed dynamics and initial state reproduced, and the reported cost equal to the objective of
    mpc.py:159-187 re-evaluated (vectorised numpy) on the returned trajectories."""
    from junction_mpc.config import PARAM_INDEX as PI
    B, T = w["B"], w["T"]
    assert (out.status == 0).all(), np.unique(out.status, return_counts=True)
    prm = w["params"] if w["params"] is not None else np.repeat(default_vector_fn(w)[None, :], B, axis=0)
    col = lambda k: prm[:, PI[k]][:, None]                 # noqa: E731
    dt = col("dt")
    tol = 1e-7
    assert (out.oa <= col("max_accel") + tol).all() and (out.oa >= col("max_decel") - tol).all()
    assert (np.abs(out.od) <= col("max_steer") + tol).all()
    assert (np.abs(np.diff(out.od, axis=1)) <= col("max_dsteer") * dt + tol).all()
    assert (out.ov <= col("speed") + tol).all() and (out.ov >= col("min_speed") - tol).all()
    np.testing.assert_allclose(out.ov[:, 1:], out.ov[:, :1] + dt * np.cumsum(out.oa, axis=1), atol=1e-9)
    np.testing.assert_allclose(np.stack([out.ox[:, 0], out.oy[:, 0], out.ov[:, 0], out.oyaw[:, 0]], 1), w["state"],
                               atol=1e-12)
    course = w["courses"][0]
    idx_last = np.minimum(w["course_len"], len(course)) - 1
    end_xy = course[idx_last, :2]
    reach = (out.xref[:, 0, :] == end_xy[:, :1]) & (out.xref[:, 1, :] == end_xy[:, 1:])
    psi = out.xref[:, 3, :]
    ex, ey = out.xref[:, 0] - out.ox, out.xref[:, 1] - out.oy
    c1, s1 = np.cos(psi + 0.5 * np.pi), np.sin(psi + 0.5 * np.pi)
    c2, s2 = np.cos(psi), np.sin(psi)
    track = col("w_perp") * (c1 * ex + s1 * ey) ** 2 + col("w_para") * (c2 * ex + s2 * ey) ** 2 \
        + col("Q_v") * out.ov ** 2 + col("Q_yaw") * (psi - out.oyaw) ** 2
    final = col("Qf_x") * ex ** 2 + col("Qf_y") * ey ** 2 + col("Qf_v") * out.ov ** 2 + col("Qf_yaw") * (psi - out.oyaw) ** 2
    stage = np.where(reach, final, track)[:, 1:].sum(1)
    ra = np.where(reach[:, :T], col("Rend_a"), col("R_a"))
    rd = np.where(reach[:, :T], col("Rend_d"), col("R_d"))
    inp = (ra * out.oa ** 2 + rd * out.od ** 2).sum(1)
    rate = (col("Rd_a") * np.diff(out.oa, axis=1) ** 2 + col("Rd_d") * np.diff(out.od, axis=1) ** 2).sum(1)
    np.testing.assert_allclose(out.cost, stage + inp + rate, rtol=1e-9)


def test_full_config2_properties(jm):
    """The whole 4096 x T=20 batch of BASELINE.json configs[1]."""
    synth, BatchedMPC = jm
    w = synth.make_workload(2)
    mpc, out = _run(BatchedMPC, w)
    _check_properties(w, out, default_vector)


@pytest.mark.parametrize("config", [3, 4])
def test_full_size_roundabout_and_multilane_properties(jm, config):
    """BASELINE.json configs[2] (65 536 instances) and configs[3] (262 144 instances) at full size, with the
    truncated course lengths coming from the collision-flag kernel as in the scenario loop."""
    synth, BatchedMPC = jm
    from oracle import collision_oracle as C
    w = synth.make_workload(config)
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=w["B"])
    margin = C.cutoff_margin(C.CarGeometry(), w["dl"])
    flag, clen = mpc.collision_host(w["agent_idx"], w["state"][:, 2], w["obstacles"], frame_window=w["frame_window"],
                                    margin=margin)
    assert 0.05 < flag.mean() < 0.95
    assert (clen[flag == 0] == len(w["courses"][0])).all() and (clen[flag == 1] > w["agent_idx"][flag == 1]).all()
    # the step searches the course from the ego index, as mpc.target_ind does in the closed loop
    w["course_len"] = clen
    w["target_ind"] = np.minimum(w["target_ind"], clen - 1).astype(np.int32)
    out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=clen)
    ok = out.status != 3          # a cut right behind the search start can leave the index rule undefined
    assert ok.mean() > 0.999
    sub = dict(w, B=int(ok.sum()), state=w["state"][ok], course_len=clen[ok], params=None)
    import types
    sel = types.SimpleNamespace(**{k: getattr(out, k)[ok] for k in ["oa", "od", "ox", "oy", "ov", "oyaw", "xref", "cost", "status"]})
    _check_properties(sub, sel, default_vector)


@pytest.mark.parametrize("T", [8, 25])
def test_full_size_sweep_properties(jm, T):
    """BASELINE.json configs[4]: one horizon slice (262 144 instances) of the 1M-instance parameter sweep."""
    synth, BatchedMPC = jm
    w = synth.make_sweep(T, states_per_point=32)
    mpc, out = _run(BatchedMPC, w)
    _check_properties(w, out, default_vector)
    assert out.iters.max() <= 40


def test_pinned_host_outputs_equal_plain_host_path(jm):
    synth, BatchedMPC = jm
    w = synth.make_workload(2, B=257)
    mpc, plain = _run(BatchedMPC, w)
    out = mpc.host_outputs(257)
    for _ in range(2):      # reuse of the same page-locked arrays
        got = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], out=out)
    assert got is out
    for key in ["oa", "od", "ox", "oy", "ov", "oyaw", "xref", "cost", "status", "iters", "target_ind", "record"]:
        assert np.array_equal(getattr(got, key), getattr(plain, key)), key
    # the packed record is what the multi-GPU all-gather ships
    assert np.array_equal(got.record[:, 0], got.od[:, 0]) and np.array_equal(got.record[:, 1], got.oa[:, 0])
    assert np.array_equal(got.record[:, 3], got.status) and np.array_equal(got.record[:, 4], got.target_ind)


@pytest.mark.parametrize("mode", ["staged", "zero_copy_results", "zero_copy"])
def test_host_transfer_modes_agree(jm, mode):
    """jmpc_step_host[_io]: staged DMA, zero-copy results (the default), zero-copy both ways give the same
    arrays, for pageable and page-locked callers, including instances that are not solved (in-out values kept)."""
    synth, BatchedMPC = jm
    w = synth.make_workload(2, B=130)
    state = w["state"].copy()
    state[3, 2] = 30.0            # v0 above the speed cap -> infeasible (status 2)
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=256)
    mpc.set_host_transfer("staged")
    ref = mpc.step_host(state, w["target_ind"], w["oa"], w["od"], course_len=w["course_len"])
    assert ref.status[3] == 2 and np.array_equal(ref.oa[3], w["oa"][3])
    mpc.set_host_transfer(mode)
    keys = ["oa", "od", "cost", "status", "iters", "target_ind", "record"]
    solved = ref.status == 0
    def same(got):
        for key in keys:
            assert np.array_equal(getattr(got, key), getattr(ref, key), equal_nan=True), (mode, key)
        for key in ["ox", "oy", "ov", "oyaw"]:           # only defined for solved instances
            assert np.array_equal(getattr(got, key)[solved], getattr(ref, key)[solved]), (mode, key)
        assert np.array_equal(got.xref, ref.xref), mode
    same(mpc.step_host(state, w["target_ind"], w["oa"], w["od"], course_len=w["course_len"]))          # pageable
    out = mpc.host_outputs(130)
    same(mpc.step_host(state, w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], out=out))   # pinned results
    pin = {k: mpc.pinned_empty(v.shape, v.dtype) for k, v in
           dict(state=state, target_ind=w["target_ind"], oa=w["oa"], od=w["od"], course_len=w["course_len"]).items()}
    for k, v in dict(state=state, target_ind=w["target_ind"], oa=w["oa"], od=w["od"], course_len=w["course_len"]).items():
        pin[k][...] = v
    same(mpc.step_host(pin["state"], pin["target_ind"], pin["oa"], pin["od"], course_len=pin["course_len"], out=out))
    assert np.array_equal(pin["oa"], w["oa"]) and np.array_equal(pin["target_ind"], w["target_ind"])     # inputs only read
    # closed loop on one set of arrays: the previous results are the next inputs, nothing is copied
    nxt_ref = mpc.step_host(state, ref.target_ind, ref.oa, ref.od, course_len=w["course_len"])
    got = mpc.step_host(pin["state"], out.target_ind, out.oa, out.od, course_len=pin["course_len"], out=out)
    for key in keys:
        assert np.array_equal(getattr(got, key), getattr(nxt_ref, key), equal_nan=True), (mode, key)
    mpc.close()


def test_results_do_not_depend_on_the_work_queue_order(jm):
    """jmpc_set_schedule: index order, a-priori key, previous-step iteration counts -- bit-identical results; the
    batch is larger than the resident warps so that the ordering kernel actually runs."""
    synth, BatchedMPC = jm
    w = synth.make_workload(2, B=4096)
    ref = None
    for mode in ["index", "apriori", "history", "history"]:         # the second "history" step runs on recorded counts
        if mode != "history" or ref is None or mpc.schedule != "history":
            mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=4096, schedule=mode)
            launches0 = mpc.launch_count
        out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"])
        assert (out.status == 0).all()
        if ref is None:
            ref = out
            continue
        for key in ["oa", "od", "ox", "oy", "ov", "oyaw", "xref", "cost", "status", "iters", "target_ind", "record"]:
            assert np.array_equal(getattr(out, key), getattr(ref, key)), (mode, key)
    assert mpc.launch_count - launches0 == 6             # (two ordering kernels + step kernel) x 2 steps on the last engine
    mpc.reset_schedule_hints()


def test_fused_record_stores_reach_the_peer_tables(jm):
    """The kernel-epilogue all-gather on one GPU: two 'peer' tables that both live on this device."""
    import torch
    synth, BatchedMPC = jm
    w = synth.make_workload(2, B=200)
    mpc, host = _run(BatchedMPC, w)
    dev = torch.device("cuda", 0)
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
    tables = [torch.full((3 * 200, 8), -7.0, dtype=torch.float64, device=dev) for _ in range(2)]
    mpc.set_record_peers([tb.data_ptr() for tb in tables], 200)          # this "rank" owns rows 200..399
    out = mpc.step(t(w["state"], torch.float64), t(w["target_ind"], torch.int32), t(w["oa"], torch.float64),
                   t(w["od"], torch.float64), mpc.alloc_outputs(200), course_len=t(w["course_len"], torch.int32))
    torch.cuda.synchronize()
    for tb in tables:
        got = tb.cpu().numpy()
        assert np.array_equal(got[200:400], host.record) and np.array_equal(got[200:400], out.record.cpu().numpy())
        assert (got[:200] == -7.0).all() and (got[400:] == -7.0).all()
    mpc.set_record_peers([], 0)


@pytest.mark.parametrize("T", [5, 10, 31])
def test_generic_horizons(jm, T):
    """Horizons without a compile-time specialisation run the generic kernel; T = 31 is the limit of the
    lane-per-stage mapping (T + 1 = 32 horizon points)."""
    synth, BatchedMPC = jm
    from junction_mpc.config import MPCConfig
    rng = np.random.default_rng(T)
    course = synth.load_course("roundabout")
    w = synth.make_states(rng, course, 40, T)
    w.update(T=T, courses=[course], params=None, B=40, dl=float(np.linalg.norm(course[0, :2] - course[1, :2])))
    mpc = BatchedMPC([course], dl=w["dl"], T=T, max_batch=64, max_T=31)
    out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"])
    refs = oracle_batch(w, range(40))
    assert compare_step(out, refs, range(40)) <= 1.0


def test_iteration_cap_is_a_failed_solve(jm):
    """max_solver_iters = 2: nothing converges.  Like any solver status other than OPTIMAL / OPTIMAL_INACCURATE in the
    reference (mpc.py:199-209, 298-301) that is a failed solve: controls untouched, xref / target reported, the
    record brakes with MAX_DECEL, and the drop-in keeps di, sets ai = MAX_DECEL and drops its warm start."""
    synth, BatchedMPC = jm
    for T, cfgno in [(20, 2), (13, 3)]:
        w = synth.make_workload(cfgno, B=64)
        mpc, out = _run(BatchedMPC, w, max_solver_iters=2)
        good = _run(BatchedMPC, w)[1]
        assert (out.status == O.STATUS_MAX_ITER).all() and (out.iters == 2).all()
        assert np.array_equal(out.oa, w["oa"]) and np.array_equal(out.od, w["od"])
        assert np.array_equal(out.xref, good.xref) and np.array_equal(out.target_ind, good.target_ind)
        assert np.isnan(out.cost).all() and np.isnan(out.record[:, 0]).all()
        assert (out.record[:, 1] == -10.0).all() and (out.record[:, 3] == O.STATUS_MAX_ITER).all()


def test_du_th_exit_of_the_linearisation_loop(jm):
    """The exit the reference left commented out (mpc.py:236-240), as an option: MAX_ITER = 4 with DU_TH."""
    synth, BatchedMPC = jm
    for cfgno, du_th in [(2, 15.0), (3, 10.0)]:      # thresholds inside the spread of du on these batches
        w = synth.make_workload(cfgno, B=48)
        w["course_len"][:] = len(w["courses"][0])
        mpc, out = _run(BatchedMPC, w, linearisation_iters=4, du_th=du_th)
        base = default_vector(w)
        refs = []
        for k in range(48):
            p = params_from_vector(base, w["T"], 4)
            c = w["courses"][0]
            refs.append(O.mpc_step(p, w["state"][k], w["oa"][k], w["od"][k], c[:, 0], c[:, 1], c[:, 2],
                                   int(w["target_ind"][k]), du_th=du_th))
        assert compare_step(out, refs, range(48)) <= 1.0
        full = _run(BatchedMPC, w, linearisation_iters=4)[1]
        assert (out.iters <= full.iters).all() and (out.iters < full.iters).any()      # some instances left early


def test_arbitrary_speed_profile_table(jm):
    """xref[2] = cv[idx] for a general per-course speed profile (mpc_with_speed.py:104), with Q_v > 0."""
    synth, BatchedMPC = jm
    from junction_mpc.config import MPCConfig, PARAM_INDEX as PI
    w = synth.make_workload(3, B=64)
    course = w["courses"][0]
    cv = 3.0 + 2.0 * np.sin(np.arange(len(course)) / 40.0)
    mpc = BatchedMPC(w["courses"], dl=w["dl"], T=w["T"], max_batch=64)
    mpc.set_course_speed([cv])
    prm = np.repeat(mpc.default_params[None, :], 64, axis=0)
    prm[:, PI["Q_v"]] = 20.0
    prm[:, PI["v_ref_cut"]] = np.where(np.arange(64) % 2 == 0, 1e9, w["target_ind"] + 25)    # half with a cut index too
    out = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=prm)
    refs = []
    for k in range(64):
        p = params_from_vector(prm[k], w["T"])
        n = int(w["course_len"][k])
        refs.append(O.mpc_step(p, w["state"][k], w["oa"][k], w["od"][k], course[:n, 0], course[:n, 1], course[:n, 2],
                               int(w["target_ind"][k]), cv=cv[:n]))
    assert compare_step(out, refs, range(64)) <= 1.0
    assert np.abs(out.xref[:, 2]).max() > 1.0
    mpc.set_course_speed(None)
    out0 = mpc.step_host(w["state"], w["target_ind"], w["oa"], w["od"], course_len=w["course_len"], params=prm)
    assert (out0.xref[:, 2] == 0).all()


def test_half_warp_pairs_are_independent(jm):
    """T <= 15 runs two instances per warp.  An instance's results must not depend on its partner: odd batch sizes,
    a failing partner (index rule / infeasible), and any permutation of the batch give identical bits."""
    synth, BatchedMPC = jm
    w = synth.make_workload(3, B=257)
    st = w["state"].copy()
    st[10, 2] = 30.0                       # infeasible partner of instance 11
    st[20, :2] += 500.0                    # far away: whatever the index rule says
    w["state"] = st
    mpc, ref = _run(BatchedMPC, w)
    assert ref.status[10] == O.STATUS_INFEASIBLE and (ref.status == 0).sum() >= 250
    perm = np.random.default_rng(0).permutation(257)
    out = mpc.step_host(st[perm], w["target_ind"][perm], w["oa"][perm], w["od"][perm], course_len=w["course_len"][perm])
    solved = ref.status == 0
    for key in ["oa", "od", "cost", "status", "iters", "target_ind", "xref"]:
        assert np.array_equal(getattr(out, key), getattr(ref, key)[perm], equal_nan=True), key
    for key in ["ox", "oy", "ov", "oyaw"]:
        assert np.array_equal(getattr(out, key)[solved[perm]], getattr(ref, key)[perm][solved[perm]]), key
    refs = oracle_batch(w, range(0, 257, 4))
    compare_step(ref, refs, range(0, 257, 4))


@pytest.mark.parametrize("config,T", [(2, 20), (3, 13), (5, 8), (5, 25), (5, 10)])
def test_low_latency_kernels_give_the_same_bits(jm, config, T):
    """Small launches (at most eight warps per SM) run the low-latency kernels: one warp per block, full register
    budget, triangular sweeps with the vector in registers (jmpc_linalg.cuh).  They must reproduce the throughput
    kernels bit for bit, so a result never depends on how many instances share its launch: 48 instances alone
    (low-latency kernel), the same 48 inside a batch of 6144 (throughput kernel), and alone on an engine with a fixed
    warp count (which switches the low-latency path off)."""
    synth, BatchedMPC = jm
    w = synth.make_workload(config, B=6144) if config != 5 else synth.make_sweep_sample(T, 6144)
    assert w["T"] == T
    small = np.arange(0, 6144, 127)[:48]
    mpc, big = _run(BatchedMPC, w)
    prm = None if w["params"] is None else w["params"][small]
    lat = mpc.step_host(w["state"][small], w["target_ind"][small], w["oa"][small], w["od"][small],
                        course_len=w["course_len"][small], params=prm)
    mpc.close()
    mpc2, fixed = _run(BatchedMPC, w, idx=small, warps_per_sm=16)
    mpc2.close()
    assert (big.status[small] == 0).sum() >= 40
    for key in ["oa", "od", "ox", "oy", "ov", "oyaw", "cost", "status", "iters", "target_ind", "xref"]:
        assert np.array_equal(getattr(lat, key), getattr(big, key)[small], equal_nan=True), key
        assert np.array_equal(getattr(lat, key), getattr(fixed, key), equal_nan=True), key
