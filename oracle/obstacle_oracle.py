"""Restatement of the reference's scripted obstacles (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows main/lib/moving_obstacles.py: ``MovingObstacleTIntersection`` (:166-232), ``MovingObstacleRoundabout``
(:28-123) and ``MovingObstacleArterial`` (:125-164), each driving a ``Bicycle`` (main/bicycle/main.py:28-41).
Pinned on tracks recorded from the reference's own classes (tests/golden/scripted_obstacles.npz, made by
tests/golden/make_golden.py --obstacles).

Quirks kept: the roundabout class overwrites its ``dt`` with 0.2 after building its Bicycle with the constructor's
dt (:45), so only the offset test sees 0.2; its ``steering_angle`` property sets ``model.theta`` as a side effect
(:82-84, :98-100) and ``get()`` reads theta before that property runs (tuple order, :118-120).
"""
from __future__ import annotations

import math

import numpy as np

KIND_TINTERSECTION, KIND_ROUNDABOUT, KIND_ARTERIAL = 1, 2, 3


def steering_angle_for_radius(radius: float, L: float = 2.86) -> float:
    """moving_obstacles.py:15-25"""
    return float(np.arctan((1 / radius) * L))


class ScriptedObstacle:
    def __init__(self, kind: int, direction: int = 1, turning: bool = False, speed: float = 25 / 3.6, offset=None,
                 dt: float = 0.2, x_init: float = 0.0, y_init: float = 0.0, initial_speed: float = 0.0, L: float = 2.86):
        self.kind, self.direction, self.turning, self.speed = kind, (1 if direction >= 0 else -1), bool(turning), speed
        self.offset = None if offset is None else offset if offset > 0 else None
        self.dt_model = dt
        self.dt = 0.2 if kind == KIND_ROUNDABOUT else dt
        self.initial_speed, self.L, self.counter = initial_speed, L, 0
        if kind == KIND_ARTERIAL:
            self.xc, self.yc, self.theta = x_init, y_init, np.pi / 2
        elif self.direction == 1:
            self.xc, self.yc, self.theta = -30, -3, 0
        else:
            self.xc, self.yc, self.theta = 30, 3, np.pi

    def steering_angle(self) -> float:
        steer = 0.0
        if not self.turning or self.kind == KIND_ARTERIAL:
            return steer
        if self.kind == KIND_TINTERSECTION:
            if self.direction == 1:
                if self.xc >= -10 and self.theta > (-np.pi / 2):
                    steer = -0.38
            elif self.xc <= 12 and self.theta < (3 * np.pi / 2):
                steer = 0.19
            return steer
        ang = steering_angle_for_radius(5)
        if self.direction == 1:
            if -7 <= self.xc <= -4 and self.yc < 0:
                steer = -ang
            if -3 < self.xc:
                steer = ang
            if self.yc > 0 and -5 <= self.xc <= -3:
                steer = -ang
            if self.xc <= -3 and self.yc > 0:
                self.theta = -np.pi
                steer = 0
        else:
            if 4 <= self.xc <= 7 and self.yc > 0:
                steer = -ang
            if self.xc < 3:
                steer = ang
            if self.yc < 0 and 3 <= self.xc <= 5:
                steer = -ang
            if 3 <= self.xc and self.yc < 0:
                self.theta = 0
                steer = 0
        return steer

    def forward_velocity(self) -> float:
        if self.offset is None or self.counter > (self.offset / self.dt):
            return self.speed
        return self.initial_speed if self.kind == KIND_ARTERIAL else 0

    def step(self):
        steer = self.steering_angle()
        v = self.forward_velocity()
        xd, yd, td = v * np.cos(self.theta), v * np.sin(self.theta), (v / self.L) * np.tan(steer)
        self.xc += xd * self.dt_model
        self.yc += yd * self.dt_model
        self.theta += td * self.dt_model
        self.counter += 1

    def get(self):
        return self.xc, self.yc, self.forward_velocity(), self.theta, 0.0, self.steering_angle()


def from_spec(row) -> ScriptedObstacle:
    """A row of tests/golden/scripted_obstacles.npz: kind, direction, turning, speed, offset (-1 = None), dt, x, y, v_init."""
    kind, direction, turning, speed, offset, dt, x0, y0, v_init = (float(v) for v in row)
    return ScriptedObstacle(int(kind), int(direction), bool(turning), speed, None if offset < 0 else offset, dt, x0, y0, v_init)
