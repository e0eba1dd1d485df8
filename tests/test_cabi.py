"""CPU-side checks of the boundary: the library loads and exports every symbol include/jmpc.h declares, the
parameter enum matches the Python table, and the host logic (config, workloads) behaves.  No compute calls."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    with open(os.path.join(ROOT, "include", "jmpc.h")) as f:
        return f.read()


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from junction_mpc import _cabi
    declared = set(re.findall(r"\b(jmpc_[a-z0-9_]+)\s*\(", _header()))
    assert declared, "no declarations found in include/jmpc.h"
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libjmpc.so does not export {name}"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    lib.jmpc_abi_version.restype = ctypes.c_int32
    lib.jmpc_nparam.restype = ctypes.c_int32
    assert lib.jmpc_abi_version() == _cabi.ABI_VERSION


def test_param_enum_matches_python_table():
    from junction_mpc.config import PARAM_NAMES, NPARAM
    body = re.search(r"enum jmpc_param \{(.*?)\};", _header(), re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n for n in re.findall(r"JMPC_(?:P_)?([A-Z_]+)", body)]
    assert names[-1] == "NPARAM"
    assert [n.lower() for n in names[:-1]] == [p.lower() for p in PARAM_NAMES]
    assert NPARAM == len(names) - 1


def test_config_derivations_follow_reference():
    from junction_mpc.config import MPCConfig, PARAM_INDEX as PI
    cfg = MPCConfig.default()
    assert cfg.T == 13 and cfg.max_iter == 1
    assert cfg.Qf == (13.0, 13.0, 0.0, 6.5)                      # Qf * T, mpc.py:28
    assert abs(cfg.max_dsteer - np.deg2rad(30.0)) == 0.0         # mpc.py:37
    v = cfg.with_T(20).param_vector(dl=0.083)
    assert v[PI["Qf_x"]] == 20.0 and v[PI["Qf_yaw"]] == 10.0
    assert v[PI["max_steer"]] == np.deg2rad(45.0) and v[PI["min_speed"]] == -5.0
    assert v[PI["sim_max_speed"]] == 30 / 3.6 and v[PI["v_ref_min"]] == 10 / 3.6


def test_reference_config_file_is_read_unchanged(tmp_path):
    import json
    from junction_mpc.config import MPCConfig, _DEFAULTS
    d = dict(_DEFAULTS)
    d.update(T=20, R=[0.1, 0.01], Rd=[10, 10])
    p = tmp_path / "mpc_config_sensitivity.json"
    p.write_text(json.dumps(d))
    cfg = MPCConfig.from_json(str(p))
    assert cfg.T == 20 and cfg.R == (0.1, 0.01) and cfg.Rd == (10.0, 10.0)


def test_workloads_are_seeded_and_valid():
    from junction_mpc import synth
    from oracle import mpc_oracle as O
    a, b = synth.make_workload(2, B=64), synth.make_workload(2, B=64)
    for k in ["state", "oa", "od", "course_len", "target_ind"]:
        assert np.array_equal(a[k], b[k])
    assert a["state"].shape == (64, 4) and a["oa"].shape == (64, 20)
    c = a["courses"][0]
    for k in range(64):        # the generator's validity filter agrees with the oracle's index rule
        n = a["course_len"][k]
        O.nearest_index_forward(a["state"][k, 0], a["state"][k, 1], c[:n, 0], c[:n, 1], int(a["target_ind"][k]))
    s = synth.make_sweep(8, states_per_point=2, max_points=16)
    from junction_mpc.config import NPARAM
    assert s["params"].shape == (32, NPARAM) and s["B"] == 32
    assert len(np.unique(s["params"], axis=0)) == 16
