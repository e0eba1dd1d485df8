"""BatchedEpisodes: B closed-loop episodes of the reference's scenario loop, device resident.

One iteration = the body of `for i in itertools.count()` in main/scenarios/mpc_intersection.py:99-163:
goal test -> ego index on the full course -> collision flag / cut -> MPC step -> plant step + history.
Everything runs in libjmpc.so kernels on torch-owned device tensors; the host only sequences launches and looks
at the `done` flags every few iterations.  Obstacles are the reference's scripted vehicles stepped on the device
(`obstacle_program`, see `scripted_obstacles`), constant-input vehicles advanced on the device (`obstacles`), or a
per-step recording (`obstacle_script`).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _cabi
from .batched import BatchedMPC


OBS_KINDS = {"constant": 0, "t_intersection": 1, "roundabout": 2, "arterial": 3}


def scripted_obstacles(specs, L: float = 2.86):
    """Device-side description of the reference's scripted obstacles (main/lib/moving_obstacles.py).

    specs: nested list [B][n_obs] of dicts with the constructor arguments of the reference classes:
      kind = "t_intersection" | "roundabout" | "arterial", direction, turning, speed, offset, dt, and for the arterial
      x_init, y_init, initial_speed.
    Returns (script [B, n_obs, 8], model [B, n_obs, 4]) float64 arrays for `BatchedEpisodes(obstacle_program=...)`:
    the constant script rows (enum jmpc_obs_script in include/jmpc.h) and the initial Bicycle state + step counter
    (start poses of moving_obstacles.py:52-61, 134-137, 190-199)."""
    B, n = len(specs), len(specs[0])
    script, model = np.zeros((B, n, 8)), np.zeros((B, n, 4))
    for b, row in enumerate(specs):
        if len(row) != n:
            raise ValueError("every episode needs the same number of obstacles")
        for k, s in enumerate(row):
            kind = OBS_KINDS[s["kind"]]
            direction = 1.0 if s.get("direction", 1) >= 0 else -1.0
            dt = float(s.get("dt", 0.2))
            offset = s.get("offset")
            aux = float(np.arctan((1 / 5) * 2.86)) if kind == 2 else float(s.get("initial_speed", 0.0))
            script[b, k] = [kind, direction, 1.0 if s.get("turning", False) else 0.0, float(s.get("speed", 25 / 3.6)),
                            -1.0 if (offset is None or offset <= 0) else float(offset), dt, 0.2 if kind == 2 else dt, aux]
            if kind == 3:
                model[b, k] = [float(s["x_init"]), float(s["y_init"]), np.pi / 2, 0.0]
            elif direction > 0:
                model[b, k] = [-30.0, -3.0, 0.0, 0.0]
            else:
                model[b, k] = [30.0, 3.0, np.pi, 0.0]
    return script, model


class BatchedEpisodes:
    def __init__(self, engine: BatchedMPC, state0, course_id=None, obstacles=None, obstacle_script=None,
                 frame_window: int = 10, margin: int = 72, params=None, max_steps: int = 256,
                 record_history: bool = True, horizon_s: float = 7.0, obstacle_program=None):
        import torch
        self.torch = torch
        self.e = engine
        self.dev = torch.device("cuda", engine.device)
        f64, i32 = torch.float64, torch.int32
        t = lambda a, dt: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=self.dev)  # noqa: E731
        self.state = t(state0, f64)
        self.B = B = self.state.shape[0]
        self.T = engine.T
        self.course_id = t(course_id, i32)
        self.params = t(params, f64)
        self.obstacles = t(obstacles, f64)                       # [B, n_obs, 6], advanced on the device
        self.script = t(obstacle_script, f64)                    # [S, B, n_obs, 6], read per step
        self.program = None
        if obstacle_program is not None:                         # the reference's scripted obstacles, stepped on the device
            if obstacles is not None or obstacle_script is not None:
                raise ValueError("give one of obstacles / obstacle_script / obstacle_program")
            self.program = (t(obstacle_program[0], f64), t(obstacle_program[1], f64))
            self.obstacles = torch.zeros(B, self.program[0].shape[1], 6, dtype=f64, device=self.dev)
        if self.obstacles is not None and self.script is not None:
            raise ValueError("give either constant-input obstacles or a script")
        self.frame_window, self.margin, self.horizon_s = int(frame_window), int(margin), float(horizon_s)
        self.max_steps = int(max_steps)
        full = torch.as_tensor(engine.course_len, dtype=i32, device=self.dev)
        self.full_len = full[self.course_id.long()] if self.course_id is not None else full[0].expand(B).contiguous()
        self.course_len = self.full_len.clone()
        z = lambda: torch.zeros(B, dtype=i32, device=self.dev)   # noqa: E731
        self.agent_idx, self.target_ind, self.done, self.steps, self.warm, self.flag = z(), z(), z(), z(), z(), z()
        self.di = torch.zeros(B, dtype=f64, device=self.dev)
        self.v = torch.zeros(B, dtype=f64, device=self.dev)
        self.oa = torch.zeros(B, self.T, dtype=f64, device=self.dev)
        self.od = torch.zeros(B, self.T, dtype=f64, device=self.dev)
        self.out = engine.alloc_outputs(B)
        self.history = torch.full((self.max_steps, B, 8), float("nan"), dtype=f64, device=self.dev) if record_history else None
        self.flags = torch.zeros(self.max_steps, B, dtype=i32, device=self.dev) if record_history else None
        self.iteration = 0
        self.iter_dev = torch.zeros(1, dtype=i32, device=self.dev)     # the same counter on the device (graph replay)
        if self.program is not None:                                 # get() before the first iteration
            self._scripted_step(advance=0)

    def _scripted_step(self, advance: int):
        e = self.e
        stream = C.c_void_p(self.torch.cuda.current_stream(e.device).cuda_stream)
        _cabi.check(e._lib.jmpc_scripted_obstacle_step(
            e._h, self.B, int(self.program[0].shape[1]), self._p(self.program[0]), self._p(self.program[1]),
            self._p(self.obstacles), self._p(self.done), int(advance), stream), "jmpc_scripted_obstacle_step")

    def _p(self, ten):
        return None if ten is None else C.c_void_p(ten.data_ptr())

    def iterate(self):
        """One loop iteration for every episode that is not done.  The engine-wide skip mask (finished episodes cost
        nothing in the step / collision kernels) is installed for the duration of the launches only: the engine never
        keeps a pointer into this object's tensors between calls."""
        self.e.set_skip_mask(self.done)
        try:
            self._iterate()
        finally:
            self.e.set_skip_mask(None)

    def _iterate(self):
        e, lib, torch = self.e, self.e._lib, self.torch
        stream = C.c_void_p(torch.cuda.current_stream(e.device).cuda_stream)
        i = self.iteration
        cfg = e.config
        _cabi.check(lib.jmpc_episode_pre(e._h, self.B, self._p(self.state), self._p(self.course_id), self._p(self.course_len),
                                         self._p(self.target_ind), self._p(self.steps), self._p(self.agent_idx),
                                         self._p(self.done), float(cfg.goal_dis), float(cfg.stop_speed), stream),
                    "jmpc_episode_pre")
        obs = self.obstacles if self.script is None else self.script[min(i, self.script.shape[0] - 1)]
        if obs is not None and obs.shape[1] > 0:
            self.v.copy_(self.state[:, 2])
            e.collision(self.agent_idx, self.v, obs, self.frame_window, self.margin, self.flag, self.course_len,
                        course_id=self.course_id, params=self.params, horizon_s=self.horizon_s)
            has_flags = True
        else:
            has_flags = False
        e.step(self.state, self.target_ind, self.oa, self.od, self.out, course_id=self.course_id,
               course_len=self.course_len, warm=self.warm, params=self.params)
        # history row, flag copy and time stamp are indexed by the device-side iteration counter: nothing in the loop
        # body depends on a host value that changes from one iteration to the next (except a script's frame)
        _cabi.check(lib.jmpc_episode_post_dev(
            e._h, self.B, self._p(self.state), self._p(self.course_id), self._p(self.out.record), self._p(self.params),
            self._p(self.target_ind), self._p(self.steps), self._p(self.done), self._p(self.di), self._p(self.warm),
            self._p(self.history), self.max_steps, self._p(self.flags) if has_flags else None, self._p(self.flag),
            self._p(self.iter_dev), float(e.dt), stream), "jmpc_episode_post_dev")
        if self.program is not None:
            self._scripted_step(advance=1)
        elif self.obstacles is not None and self.obstacles.shape[1] > 0:
            _cabi.check(lib.jmpc_obstacle_step(e._h, self.B, int(self.obstacles.shape[1]), self._p(self.obstacles),
                                               self._p(self.done), float(e.dt), stream), "jmpc_obstacle_step")
        _cabi.check(lib.jmpc_counter_add(e._h, self._p(self.iter_dev), 1, stream), "jmpc_counter_add")
        self.iteration += 1

    def run(self, max_steps: Optional[int] = None, check_every: int = 8, use_graph: bool = False):
        """Iterate until every episode reached its goal (or max_steps).  Returns a dict of numpy results.

        With constant-input obstacles the loop body takes no per-iteration host argument, so with `use_graph`
        `check_every` iterations are captured once as a CUDA graph and replayed (one launch per 8 iterations instead
        of ~50 driver calls).  Measured on 4096 episodes (T = 13, ~290 iterations, 0.164 s): the capture costs more
        than the replays save (0.167-0.183 s), the loop is not launch bound at this size, so it is off by default;
        it pays for long runs of small batches."""
        torch = self.torch
        max_steps = int(max_steps or self.max_steps)
        graph = None
        if use_graph and self.script is None and max_steps - self.iteration > 2 * check_every:
            self.iterate()                                    # eager once: scratch allocations, kernel attributes
            torch.cuda.synchronize(self.e.device)
            it0 = self.iteration
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="relaxed"):
                    for _ in range(check_every):
                        self.iterate()
                self.iteration = it0                          # capturing executed nothing
            except Exception:                                 # capture refused: plain launches
                graph = None
                self.iteration = it0
                torch.cuda.synchronize(self.e.device)
        while self.iteration < max_steps:
            if graph is not None and max_steps - self.iteration >= check_every:
                graph.replay()
                self.iteration += check_every
                if bool((self.done != 0).all().item()):
                    break
                continue
            self.iterate()
            if self.iteration % check_every == 0 and bool((self.done != 0).all().item()):
                break
        # the reference tests the goal at the top of the next iteration
        lib, e = self.e._lib, self.e
        stream = C.c_void_p(self.torch.cuda.current_stream(e.device).cuda_stream)
        _cabi.check(lib.jmpc_episode_pre(e._h, self.B, self._p(self.state), self._p(self.course_id), self._p(self.course_len),
                                         self._p(self.target_ind), self._p(self.steps), self._p(self.agent_idx),
                                         self._p(self.done), float(e.config.goal_dis), float(e.config.stop_speed), stream),
                    "jmpc_episode_pre")
        self.torch.cuda.synchronize(e.device)
        res = dict(steps=self.steps.cpu().numpy(), done=self.done.cpu().numpy(), state=self.state.cpu().numpy(),
                   iterations=self.iteration)
        if self.history is not None:
            n = min(self.iteration, self.max_steps)
            # [steps, B, 8]: x, y, yaw, v, t, delta, a, xref_dev; row i is what History.store appends after step i
            # (t = (i + 2) dt: HistorySimulation stores the initial state first, at t = dt; simulation.py:53-61)
            res["history"] = self.history[:n].cpu().numpy()
            res["flags"] = self.flags[:n].cpu().numpy()
        return res
