"""MPC parameters: the reference's JSON keys, their derivations, and the per-instance parameter vector.

The JSON schema is the reference's `main/config/mpc_config.json` (read at import time by
`main/lib/mpc.py:14-39`) and is consumed unchanged: point `MPCConfig.from_json` (or the environment variable
`JMPC_CONFIG`) at that file.  Without a file the built-in defaults below equal the reference's default file.

Derivations reproduced from mpc.py: `Qf = diag(Qf) * T` (:28), `MAX_DSTEER = deg2rad(MAX_DSTEER)` (:37); class
constants from `main/lib/simulation.py:23-25` (MAX_STEER = 45 deg, MAX_SPEED = 30/3.6, MIN_SPEED = -5).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, replace
from typing import Optional, Tuple

import numpy as np

# Row order of the per-instance parameter block `params[NPARAM, B]` of the C ABI (include/jmpc.h keeps the
# same enum).  Every row can differ per instance, which is what the sensitivity sweep needs.
PARAM_NAMES = (
    "dt", "dl", "L", "speed", "w_perp", "w_para", "R_a", "R_d", "Rd_a", "Rd_d", "Q_v", "Q_yaw",
    "Qf_x", "Qf_y", "Qf_v", "Qf_yaw", "Rend_a", "Rend_d", "max_dsteer", "max_accel", "max_decel",
    "max_steer", "sim_max_speed", "min_speed", "v_ref_min", "v_ref", "v_ref_cut",
)
PARAM_INDEX = {k: i for i, k in enumerate(PARAM_NAMES)}
NPARAM = len(PARAM_NAMES)

_DEFAULTS = {
    "NX": 4, "NU": 2, "T": 13, "w_perp": 20.0, "w_para": 1.0, "R": [0.01, 0.01], "Rd": [0.01, 1.0],
    "Q_v_yaw": [0.0, 0.5], "Qf": [1.0, 1.0, 0.0, 0.5], "GOAL_DIS": 1.5, "STOP_SPEED": 0.1389, "MAX_TIME": 13.0,
    "MAX_ITER": 1, "DU_TH": 0.1, "MAX_DSTEER": 30.0, "MAX_ACCEL": 2.0, "MAX_DECEL": -10,
}

SIM_MAX_STEER = float(np.deg2rad(45.0))
SIM_MAX_SPEED = 30.0 / 3.6
SIM_MIN_SPEED = -5.0
V_REF_MIN = 10.0 / 3.6           # mpc.py:99
R_END = (10.0, 10.0)             # mpc.py:181


@dataclass(frozen=True)
class MPCConfig:
    T: int
    w_perp: float
    w_para: float
    R: Tuple[float, float]
    Rd: Tuple[float, float]
    Q_v_yaw: Tuple[float, float]
    Qf_raw: Tuple[float, float, float, float]       # as in the file, BEFORE the * T of mpc.py:28
    goal_dis: float
    stop_speed: float
    max_time: float
    max_iter: int
    du_th: float
    max_dsteer_deg: float
    max_accel: float
    max_decel: float
    nx: int = 4
    nu: int = 2

    @staticmethod
    def from_dict(d: dict) -> "MPCConfig":
        if int(d.get("NX", 4)) != 4 or int(d.get("NU", 2)) != 2:
            raise ValueError("the kinematic-bicycle MPC has NX=4, NU=2")
        return MPCConfig(
            T=int(d["T"]), w_perp=float(d["w_perp"]), w_para=float(d["w_para"]),
            R=(float(d["R"][0]), float(d["R"][1])), Rd=(float(d["Rd"][0]), float(d["Rd"][1])),
            Q_v_yaw=(float(d["Q_v_yaw"][0]), float(d["Q_v_yaw"][1])),
            Qf_raw=tuple(float(v) for v in d["Qf"]), goal_dis=float(d["GOAL_DIS"]),
            stop_speed=float(d["STOP_SPEED"]), max_time=float(d["MAX_TIME"]), max_iter=int(d["MAX_ITER"]),
            du_th=float(d["DU_TH"]), max_dsteer_deg=float(d["MAX_DSTEER"]), max_accel=float(d["MAX_ACCEL"]),
            max_decel=float(d["MAX_DECEL"]))

    @staticmethod
    def from_json(path: str) -> "MPCConfig":
        with open(path, "r") as f:
            return MPCConfig.from_dict(json.load(f))

    @staticmethod
    def default() -> "MPCConfig":
        path = os.environ.get("JMPC_CONFIG")
        return MPCConfig.from_json(path) if path else MPCConfig.from_dict(_DEFAULTS)

    def with_T(self, T: int) -> "MPCConfig":
        return replace(self, T=int(T))

    @property
    def Qf(self) -> Tuple[float, float, float, float]:
        return tuple(v * self.T for v in self.Qf_raw)

    @property
    def max_dsteer(self) -> float:
        return float(np.deg2rad(self.max_dsteer_deg))

    def param_vector(self, dl: float, dt: float = 0.2, L: float = 2.86, speed: float = 30.0 / 3.6) -> np.ndarray:
        v = np.zeros(NPARAM)
        qf = self.Qf
        for key, val in [
            ("dt", dt), ("dl", dl), ("L", L), ("speed", speed), ("w_perp", self.w_perp), ("w_para", self.w_para),
            ("R_a", self.R[0]), ("R_d", self.R[1]), ("Rd_a", self.Rd[0]), ("Rd_d", self.Rd[1]),
            ("Q_v", self.Q_v_yaw[0]), ("Q_yaw", self.Q_v_yaw[1]), ("Qf_x", qf[0]), ("Qf_y", qf[1]),
            ("Qf_v", qf[2]), ("Qf_yaw", qf[3]), ("Rend_a", R_END[0]), ("Rend_d", R_END[1]),
            ("max_dsteer", self.max_dsteer), ("max_accel", self.max_accel), ("max_decel", self.max_decel),
            ("max_steer", SIM_MAX_STEER), ("sim_max_speed", SIM_MAX_SPEED), ("min_speed", SIM_MIN_SPEED),
            ("v_ref_min", V_REF_MIN), ("v_ref", 0.0), ("v_ref_cut", 1e9),
        ]:
            v[PARAM_INDEX[key]] = val
        return v
