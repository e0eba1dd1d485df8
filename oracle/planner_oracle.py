"""Restatement of the reference's motion-primitive A* planner on plain arrays (TEST INFRASTRUCTURE ONLY, see
oracle/__init__.py).

Follows main/lib/a_star.py:31-78 (the search), main/lib/mp_search_ww_generic.py:27-256 (neighbours, costs, heuristic,
goal test, trajectory assembly), main/lib/obstacles.py:157-176 (half-plane collision test), main/lib/linalg.py (the 2-D
transform), main/lib/maths.py (angle normalisation) and main/lib/trajectories.py:10-86 (collision-check points of a
primitive).  Pinned on searches recorded from the reference's own classes (tests/golden/planner.npz, made by
tests/golden/make_golden.py --planner): expansion order, node path, primitive per edge, cost, full trajectory.

Inputs are what the reference objects reduce to: start / goal tuples, the goal box, per-obstacle half-plane rows
(`Obstacle.to_convex(margin)`), nine weights, and the primitive set (points, total length).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from heapq import heappop, heappush
from typing import List, Optional, Sequence

import numpy as np

WEIGHT_KEYS = ["wh_dist", "wh_theta", "wh_steering", "wh_obstacle", "wh_center", "wc_dist", "wc_steering", "wc_obstacle",
               "wc_center"]
DEFAULT_WEIGHTS = np.array([1.0, 2.7, 15.0, 0.0, 0.0, 1.0, 5.0, 0.1, 0.0])     # mp_search_ww_generic.py:27-31

STATUS_FOUND, STATUS_NO_SOLUTION, STATUS_LIMIT = 0, 1, 2


def normalize_angle(theta: float) -> float:
    """maths.py"""
    theta = theta % math.tau
    if theta >= math.pi:
        theta -= math.tau
    return theta


def transform_points(node, points: np.ndarray) -> np.ndarray:
    """linalg.py: create_2d_transform_mtx + transform_2d_pts for (N, 3) points [x, y, theta]."""
    x, y, theta = node
    c, s = np.cos(theta), np.sin(theta)
    xy = points[:, :2]
    if x == 0 and y == 0:
        out = xy @ np.array([[c, -s], [s, c]]).T
    else:
        mtx = np.array([[c, -s, x], [s, c, y], [0, 0, 1]])
        out = (np.append(xy, np.ones((len(points), 1)), axis=1) @ mtx.T)[:, :2]
    return np.append(out, points[:, 2:] + theta, axis=1)


def resample_curve(points: np.ndarray, dl: float) -> np.ndarray:
    """trajectories.py:58-86 with keep_last_point=True."""
    step = np.append(0.0, np.linalg.norm(points[1:, :2] - points[:-1, :2], axis=1))
    bucket = np.floor(step.cumsum() / dl).astype(int)
    mask = np.append(True, (bucket[1:] - bucket[:-1]) >= 1.0)
    mask[-1] = True
    return points[mask].copy()


def collision_points(mp_points: np.ndarray, radius: float, circle_centers: np.ndarray) -> np.ndarray:
    """mp_search_ww_generic.py:121-138: resample at the car radius, then one point per collision circle
    (trajectories.py:10-55), concatenated circle by circle."""
    pts = resample_curve(mp_points, radius)
    th = pts[:, 2]
    out = []
    for cx, cy in circle_centers:
        off = np.vstack([np.cos(th) * cx - np.sin(th) * cy, np.sin(th) * cx + np.cos(th) * cy]).T
        off += pts[:, :2]
        out.append(np.append(off, np.atleast_2d(th).T, axis=1))
    return np.concatenate(out, axis=0)


@dataclass
class PlanResult:
    status: int
    cost: float
    path: np.ndarray            # (n, 3) nodes
    mp_idx: np.ndarray          # (n - 1,) primitive of every edge
    trajectory: np.ndarray      # (m, 3)
    expanded: np.ndarray        # (k, 5): g, h, x, y, theta in expansion order


class Planner:
    def __init__(self, mp_points: np.ndarray, mp_total_length: Sequence[float], radius: float, circle_centers: np.ndarray):
        self.mp_points = np.asarray(mp_points, float)                # [n_mp, n_pts, 3]
        self.mp_len = [float(v) for v in mp_total_length]
        self.cc = [collision_points(p, radius, np.asarray(circle_centers, float)) for p in self.mp_points]

    def plan(self, start, goal_point, goal_area, allowed_dtheta: float, hp: np.ndarray, hp_n: np.ndarray,
             weights: Optional[np.ndarray] = None, max_expansions: int = 100000) -> PlanResult:
        w = dict(zip(WEIGHT_KEYS, DEFAULT_WEIGHTS if weights is None else [float(v) for v in weights]))
        start = tuple(float(v) for v in start)
        gx, gy, gth = (float(v) for v in goal_point)
        x1, y1, x2, y2 = (float(v) for v in goal_area)
        obstacles = [np.asarray(hp[k, :hp_n[k]], float) for k in range(len(hp_n))]
        edge_mp = {}

        def steer_cost(a_theta, b_theta):
            d = b_theta - a_theta
            d = (d + np.pi) % (2 * np.pi) - np.pi
            return abs(d)

        def nearest_obstacle(x, y):
            best = float("inf")
            for o in obstacles:
                d = min(abs(a * x + b * y + c) / (a ** 2 + b ** 2) ** 0.5 for a, b, c in o)
                if d < best:
                    best = d
            return best

        def heuristic(node):
            x, y, theta = node
            dxy = np.sqrt((x - gx) ** 2 + (y - gy) ** 2)
            dth = min(abs(theta - gth), abs(theta - gth) - allowed_dtheta / 2)
            sc = steer_cost(theta, gth)
            oc = dc = 0.0
            if w["wh_obstacle"] != 0.0:
                d = nearest_obstacle(x, y)
                oc = 1 / d if d else float("inf")
            if w["wh_center"] != 0.0:
                dc = np.sqrt(x ** 2 + y ** 2)
            return w["wh_dist"] * dxy + w["wh_theta"] * dth + w["wh_steering"] * sc + w["wh_obstacle"] * oc + w["wh_center"] * dc

        def is_goal(node):
            x, y, theta = node
            dx, dy = max(x1 - x, 0, x - x2), max(y1 - y, 0, y - y2)
            return np.sqrt(dx * dx + dy * dy) <= 1e-5 and abs(theta - gth) <= allowed_dtheta

        def neighbours(node):
            for m in range(len(self.mp_points)):
                pts = transform_points(node, self.cc[m])[:, :2].T
                homog = np.vstack([pts, np.ones((pts.shape[1],))])
                if any(bool(np.any(np.all((o @ homog) <= 0, axis=0))) for o in obstacles):
                    continue
                x, y, theta = tuple(np.squeeze(transform_points(node, np.atleast_2d(self.mp_points[m][-1]))).tolist())
                nb = (x, y, normalize_angle(theta))
                edge_mp[node, nb] = m
                sc = steer_cost(node[2], nb[2])
                oc = dc = 0.0
                if w["wh_obstacle"] != 0.0:                     # (sic) the heuristic's weight switches this term on
                    d = nearest_obstacle(nb[0], nb[1])
                    oc = 1 / d if d else float("inf")
                if w["wc_center"] != 0.0:
                    dc = np.linalg.norm([x, y])
                yield (w["wc_dist"] * self.mp_len[m] + w["wc_steering"] * sc + w["wc_obstacle"] * oc + w["wc_center"] * dc), nb

        q = [(0, 0, start, start)]
        pred = {}
        log: List[list] = []
        status, cost, path = STATUS_NO_SOLUTION, float("nan"), []
        while q:
            gh, g, node, p = heappop(q)
            if node in pred and g >= pred[node][0]:
                continue
            if len(log) >= max_expansions:
                status = STATUS_LIMIT
                break
            log.append([g, gh - g, *node])
            pred[node] = g, p
            if is_goal(node):
                path = [node]
                while node != start:
                    path.append(p)
                    node, p = p, pred[p][1]
                path.reverse()
                status, cost = STATUS_FOUND, g
                break
            for edge, nb in neighbours(node):
                ng = g + edge
                if nb not in pred or ng < pred[nb][0]:
                    heappush(q, (ng + heuristic(nb), ng, nb, node))
        mp_idx = [edge_mp[a, b] for a, b in zip(path[:-1], path[1:])]
        traj = [transform_points(a, self.mp_points[m])[:-1] for a, m in zip(path[:-1], mp_idx)]
        return PlanResult(status=status, cost=float(cost), path=np.array(path, float).reshape(-1, 3),
                          mp_idx=np.array(mp_idx, int), trajectory=np.concatenate(traj, axis=0) if traj else np.zeros((0, 3)),
                          expanded=np.array(log, float).reshape(-1, 5))
