"""Float64 numpy restatement of the collision-flag producer (TEST INFRASTRUCTURE ONLY).

Follows (paths relative to /root/reference/):
  main/lib/trajectories.py:58-86                  resample_curve
  main/scenarios/mpc_intersection.py:114-143      ego max-acceleration prediction, cut index
  main/lib/moving_obstacles_prediction.py:21-47   constant-input obstacle prediction
  main/lib/collision_avoidance.py:68-124,168-180  time-shifted circle test, cut lookup
  main/lib/car_dimensions.py:62-90                collision circles of the bicycle-model car

The reference builds one big (rows x 2) table whose row order decides which obstacle circle is
reported when several pairs touch in the same frame.  The order is: frame, then ego circle, then
obstacle copy (obstacle-major, frame offset -fw..+fw minor), then obstacle circle.  This file keeps
that order with a 4-D array instead of the reference's reshape/repeat construction.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class CarGeometry:
    """BicycleModelDimensions (car_dimensions.py:82-90): width 2.0, length L + 0.64, anchor = rear axle."""
    L: float = 2.86
    width: float = 2.0
    extra_length: float = 0.64

    @property
    def radius(self) -> float:
        return self.width / (2 ** .5)

    @property
    def circle_offsets(self) -> Tuple[float, float]:
        """x offsets (object space, y offset is 0) of the front and rear circle centres."""
        length = self.L + self.extra_length
        half_gap = length / 2 - self.width / 2
        mid = self.L / 2
        return (mid + half_gap, mid - half_gap)


def circle_tracks(traj: np.ndarray, geo: CarGeometry) -> np.ndarray:
    """traj (F, >=3) [x, y, theta] -> (2, F, 2) world positions of the two circle centres.
    Same arithmetic order as trajectories.py:27-34 with y offset 0."""
    th = traj[:, 2]
    out = np.empty((2, len(traj), 2))
    for k, off in enumerate(geo.circle_offsets):
        out[k, :, 0] = (np.cos(th) * off - np.sin(th) * 0.0) + traj[:, 0]
        out[k, :, 1] = (np.sin(th) * off + np.cos(th) * 0.0) + traj[:, 1]
    return out


def resample_curve(points: np.ndarray, dl, keep_last_point: bool = True) -> np.ndarray:
    seg = np.sqrt(((points[1:, :2] - points[:-1, :2]) ** 2).sum(axis=1))
    arc = np.concatenate([[0.0], seg]).cumsum()
    bucket = np.floor(arc / dl).astype(int)
    keep = np.concatenate([[True], (bucket[1:] - bucket[:-1]) >= 1])
    if keep_last_point:
        keep[-1] = True
    return points[keep].copy()


def ego_prediction(path_from_agent: np.ndarray, v: float, dt: float, max_accel: float,
                   max_speed: float) -> np.ndarray:
    """Ego trajectory assuming it accelerates as hard as it can (mpc_intersection.py:114-120).
    Note the reference adds MAX_ACCEL per *path point*, not per second; kept as is."""
    if v < max_speed:
        dl = np.cumsum(np.zeros(path_from_agent.shape[0]) + max_accel) + v
        dl = dt * np.minimum(dl, max_speed)
        return resample_curve(path_from_agent, dl)
    return resample_curve(path_from_agent, dt * max_speed)


def predict_obstacle(x, y, v, yaw, a, steer, dt: float, L: float, horizon: float = 7.0) -> np.ndarray:
    """(F, 4) rows [x, y, yaw, t]; position uses the old speed/yaw, yaw uses the NEW speed."""
    n = len(np.arange(0, horizon, dt))
    out = np.zeros((n, 4))
    for k in range(n):
        x += v * math.cos(yaw) * dt
        y += v * math.sin(yaw) * dt
        v += a * dt
        yaw += (v / L) * math.tan(steer) * dt
        out[k] = (x, y, yaw, k * dt)
    return out


def _shift(traj: np.ndarray, off: int) -> np.ndarray:
    """Frame-shifted copy padded with the first/last sample (collision_avoidance.py:68-82)."""
    if off < 0:
        return np.concatenate([traj[-off:], np.repeat(traj[-1:], -off, axis=0)], axis=0)
    if off > 0:
        return np.concatenate([np.repeat(traj[:1], off, axis=0), traj[:-off]], axis=0)
    return traj


def _pad(traj: np.ndarray, n: int) -> np.ndarray:
    if len(traj) < n:
        return np.vstack([traj, np.repeat(traj[-1:], n - len(traj), axis=0)])
    return traj[:n]


def check_collision(geo: CarGeometry, traj_agent: np.ndarray, path_detailed: np.ndarray,
                    traj_obstacles: Sequence[np.ndarray], frame_window: int) -> Optional[Tuple[float, float, int]]:
    """Returns (x, y, index into path_detailed) of the cut point, or None."""
    if len(traj_obstacles) == 0:
        return None
    copies: List[np.ndarray] = [_shift(tr, off) for tr in traj_obstacles
                                for off in range(-frame_window, frame_window + 1)]
    frames = max(len(traj_agent), max(len(c) for c in copies))
    ego = circle_tracks(_pad(traj_agent, frames), geo)                    # (2, F, 2)
    obs = np.stack([circle_tracks(_pad(c, frames), geo) for c in copies])  # (J, 2, F, 2)
    # D[f, ego circle, copy, obstacle circle]
    diff = ego.transpose(1, 0, 2)[:, :, None, None, :] - obs.transpose(2, 0, 1, 3)[:, None, :, :, :]
    dist = np.sqrt((diff * diff).sum(axis=-1))
    reach = 2 * geo.radius
    hit = (dist <= reach).reshape(-1)
    first = int(np.argmax(hit))
    if not hit[first]:
        return None
    f, rem = divmod(first, 2 * len(copies) * 2)
    _, rem = divmod(rem, len(copies) * 2)
    j, oc = divmod(rem, 2)
    where = obs[j, oc, f]
    ego_all = circle_tracks(path_detailed, geo).reshape(-1, 2)            # front-circle rows first
    d2 = where - ego_all
    near = np.sqrt((d2 * d2).sum(axis=1)) <= reach
    k = int(np.argmax(near)) % len(path_detailed)
    return float(path_detailed[k, 0]), float(path_detailed[k, 1]), k


def cut_index(path_full: np.ndarray, x: float, y: float, radius: float = 0.001) -> int:
    d = np.sqrt((path_full[:, 0] - x) ** 2 + (path_full[:, 1] - y) ** 2) <= radius
    k = int(np.argmax(d))
    if not d[k]:
        raise ValueError("cut point not on the path (reference returns the array itself here)")
    return k


def collision_cut(geo: CarGeometry, path_full: np.ndarray, agent_idx: int, v: float,
                  obstacles: Sequence[Sequence[float]], dt: float, frame_window: int, max_accel: float,
                  max_speed: float, margin: int, horizon: float = 7.0) -> Tuple[bool, int]:
    """One flag evaluation as the scenario loop does it (mpc_intersection.py:111-140).
    obstacles: iterable of (x, y, v, yaw, a, steer).  Returns (collision flag, effective course length):
    the MPC is then fed path_full[:length]; length == len(path_full) when there is no collision."""
    path = path_full[agent_idx:]
    ego = ego_prediction(path, v, dt, max_accel, max_speed)
    preds = [predict_obstacle(*o, dt=dt, L=geo.L, horizon=horizon) for o in obstacles]
    hit = check_collision(geo, ego, path, preds, frame_window)
    if hit is None:
        return False, len(path_full)
    k = cut_index(path_full, hit[0], hit[1]) - margin
    return True, max(agent_idx + 1, k)


def cutoff_margin(geo: CarGeometry, dl: float) -> int:
    """EXTRA_CUTOFF_MARGIN (mpc_intersection.py:87-88)."""
    return 4 * int(math.ceil(geo.radius / dl))
