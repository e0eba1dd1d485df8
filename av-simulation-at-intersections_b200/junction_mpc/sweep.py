"""The reference's sensitivity study as one batched run, with its results as data files instead of plots.

Reference: main/scenarios/mpc_sensitivity_analysis_comulative.py.  The script studies one parameter at a time, five
values each, by rewriting `mpc_config_sensitivity.json` and re-running a whole closed-loop episode
(`lib.mpc_sensitivity.MPC`, no obstacles, speed cap = Simulation.MAX_SPEED), then draws speed / acceleration /
deviation / trajectory comparisons from the `History` of every run (:268-272, :319-369, :403-437) into
`results/mpc_sensitivity/<parameter>_{speed,acceleration,deviation,trajectories}.pdf`.

Here all runs of all studies are episodes of ONE batch with per-instance parameters (`BatchedEpisodes`), and the
artefacts are `<parameter>_histories.npz` (the History tables) and `<parameter>_summary.csv`.
"""
from __future__ import annotations

import csv
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from .batched import BatchedMPC
from .config import MPCConfig, PARAM_INDEX, SIM_MAX_SPEED
from .episodes import BatchedEpisodes

# `reset_config` of the reference script (:32-51): the sweep's baseline
SENSITIVITY_DEFAULTS = {
    "NX": 4, "NU": 2, "T": 13, "w_perp": 20.0, "w_para": 1.0, "R": [0.1, 0.01], "Rd": [10, 1.0],
    "Q_v_yaw": [0.0, 0.5], "Qf": [1.0, 1.0, 0.0, 0.5], "GOAL_DIS": 1.5, "STOP_SPEED": 0.1389, "MAX_TIME": 13.0,
    "MAX_ITER": 1, "DU_TH": 0.1, "MAX_DSTEER": 30.0, "MAX_ACCEL": 2.0, "MAX_DECEL": -10,
}

# study name (the script's `parameter_in_study_name_save`) -> (parameter row, values)   (:100-128)
REFERENCE_STUDIES = {
    "w_perp": ("w_perp", [0, 1, 10, 20, 50]),
    "w_para": ("w_para", [0, 0.1, 1, 5, 10]),
    "R_acc": ("R_a", [0, 0.01, 0.1, 1, 10]),
    "R_steer": ("R_d", [0.0, 0.01, 0.1, 1, 10]),
    "Rd_acc": ("Rd_a", [0, 1, 5, 10, 20]),
    "Rd_steer": ("Rd_d", [0, 0.01, 0.1, 1, 10]),
}

HISTORY_COLUMNS = ("x", "y", "yaw", "v", "t", "delta", "a", "xref_deviation")      # simulation.py:64-84


@dataclass
class SweepRun:
    study: str
    value: float
    steps: int
    goal_reached: bool
    history: np.ndarray          # [steps + 1, 8], first row = the initial state as HistorySimulation stores it


def run_sensitivity_studies(course: np.ndarray, studies: Optional[Dict[str, tuple]] = None,
                            config: Optional[MPCConfig] = None, dt: float = 0.2, L: float = 2.86,
                            max_steps: int = 400, device: int = 0) -> List[SweepRun]:
    """course: (N, 3) x, y, smoothed yaw.  Runs every value of every study as one batch of episodes."""
    studies = studies or REFERENCE_STUDIES
    cfg = config or MPCConfig.from_dict(SENSITIVITY_DEFAULTS)
    dl = float(np.linalg.norm(course[0, :2] - course[1, :2]))
    base = cfg.param_vector(dl=dl, dt=dt, L=L, speed=SIM_MAX_SPEED)          # mpc_sensitivity.py:207
    rows, labels = [], []
    for name, (key, values) in studies.items():
        for val in values:
            p = base.copy()
            p[PARAM_INDEX[key]] = float(val)
            rows.append(p)
            labels.append((name, float(val)))
    params = np.array(rows)
    B = len(rows)
    engine = BatchedMPC([course], dl=dl, T=cfg.T, config=cfg, dt=dt, L=L, speed=SIM_MAX_SPEED, max_batch=B, device=device)
    state0 = np.repeat(np.array([[course[0, 0], course[0, 1], 0.0, course[0, 2]]]), B, axis=0)
    ep = BatchedEpisodes(engine, state0, params=params, max_steps=max_steps)
    res = ep.run(max_steps=max_steps)
    runs = []
    for b, (name, val) in enumerate(labels):
        n = int(res["steps"][b])
        first = np.array([[state0[b, 0], state0[b, 1], state0[b, 3], state0[b, 2], dt, 0.0, 0.0, 0.0]])
        runs.append(SweepRun(study=name, value=val, steps=n, goal_reached=bool(res["done"][b] == 1),
                             history=np.vstack([first, res["history"][:n, b]])))
    engine.close()
    return runs


def summarise(run: SweepRun) -> dict:
    h = run.history
    dev = h[1:, 7]
    return {"study": run.study, "value": run.value, "steps": run.steps, "sim_time_s": float(h[-1, 4]),
            "goal_reached": int(run.goal_reached), "max_speed_kmh": float(h[:, 3].max() * 3.6),
            "max_abs_accel": float(np.abs(h[:, 6]).max()), "max_abs_steer": float(np.abs(h[:, 5]).max()),
            "max_xref_deviation": float(np.nanmax(dev)) if len(dev) else 0.0,
            "rms_xref_deviation": float(np.sqrt(np.nanmean(dev ** 2))) if len(dev) else 0.0}


def save_studies(runs: Sequence[SweepRun], outdir: str) -> List[str]:
    """Writes `<study>_histories.npz` (one [steps+1, 8] table per value, columns HISTORY_COLUMNS) and
    `<study>_summary.csv` for every study; returns the file names."""
    os.makedirs(outdir, exist_ok=True)
    files = []
    for name in sorted({r.study for r in runs}):
        sel = [r for r in runs if r.study == name]
        arrays = {f"value_{k}": r.history for k, r in enumerate(sel)}
        arrays["values"] = np.array([r.value for r in sel])
        arrays["columns"] = np.array(HISTORY_COLUMNS)
        path = os.path.join(outdir, f"{name}_histories.npz")
        np.savez_compressed(path, **arrays)
        files.append(path)
        path = os.path.join(outdir, f"{name}_summary.csv")
        with open(path, "w", newline="") as f:
            rows = [summarise(r) for r in sel]
            wr = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
            wr.writeheader()
            wr.writerows(rows)
        files.append(path)
    return files
