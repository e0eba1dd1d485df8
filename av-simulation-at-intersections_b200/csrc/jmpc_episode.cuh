// jmpc_episode.cuh -- the pieces of the scenario loop around the MPC step, for B closed-loop episodes on the device.
//
// Reference loop: main/scenarios/mpc_intersection.py:99-163 (identical in mpc_roundabout.py / _multi_lane.py):
//   is_goal(state) -> break                                  mpc.py:314-330            [episode_pre_kernel]
//   traj_agent_idx = nearest forward index on the full course, unless the truncated course has collapsed onto
//                    the ego's point                         mpc_intersection.py:107-109 [episode_pre_kernel]
//   collision check -> truncated course length               :111-143                  [collision_kernel]
//   delta, a = mpc.step(state)                               :146                      [mpc_step_kernel]
//   xref deviation; plant step; history                      :163, mpc.py:305-312, simulation.py:35-61 [episode_post_kernel]
// Obstacles either move with constant inputs (obstacle_step_kernel: the update of their predictor,
// moving_obstacles_prediction.py:21-29) or follow the reference's scripted steering rules
// (scripted_obstacle_kernel: MovingObstacleTIntersection / Roundabout / Arterial, moving_obstacles.py:28-232).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/jmpc.h"
#include "jmpc_collision.cuh"
#include "jmpc_step.cuh"              // nearest_index

namespace jmpc {

struct EpisodeArgs {
  int B, T;
  const double* cx; const double* cy; const double* cyaw; const int* course_n; int course_stride;
  const int* course_id;
  const double* params; ParamBlock defaults;
  double goal_dis, stop_speed;
  double* state;            // [B][4] x, y, v, yaw   (in-out)
  int* agent_idx;           // [B]   traj_agent_idx  (in-out)
  int* course_len;          // [B]   length of the course the MPC was last given (len(mpc.cx))
  int* target_ind;          // [B]   mpc.target_ind
  int* done;                // [B]   1 once is_goal fired (or the index rule failed: done = 2)
  int* steps;               // [B]   loop iterations executed
  double* di;               // [B]   last commanded steer (kept when a solve fails, mpc.py:298-301)
  int* warm;                // [B]   1 when the next step may use oa/od as its linearisation point (mpc.py:225-227)
  const double* record;     // [B][JMPC_RECORD_LEN] from the step kernel
  double* history;          // [B][8] slice for this step: x, y, yaw, v, t, delta, a, xref_deviation; or nullptr
  double t_now;
  // device-indexed variant (jmpc_episode_post_dev): the loop iteration is read from device memory, so that a
  // captured CUDA graph of the loop body can be replayed without per-iteration host arguments
  const int* iter_dev;      // nullptr -> use history / t_now above
  double* history_base;     // [rows][B][8] or nullptr
  int* flags_base;          // [rows][B] or nullptr: per-iteration copy of the collision flags
  const int* flag;          // [B]
  int history_rows;
  double dt_loop;
};

// One warp per episode: goal test, then the ego's index on the full course.
__global__ void __launch_bounds__(128) episode_pre_kernel(const EpisodeArgs A) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= A.B) return;
  if (A.done[b]) return;
  const int cid = A.course_id ? A.course_id[b] : 0;
  const double* cx = A.cx + (size_t)cid * A.course_stride;
  const double* cy = A.cy + (size_t)cid * A.course_stride;
  const double* cyaw = A.cyaw + (size_t)cid * A.course_stride;
  const int N = A.course_n[cid];
  const double x = A.state[4 * b], y = A.state[4 * b + 1], v = A.state[4 * b + 2];
  // is_goal: distance to the end of the FULL course, target index against the CURRENT course length (mpc.py:322)
  const int len = A.course_len[b];
  const double d = hypot(x - cx[N - 1], y - cy[N - 1]);
  bool goal = d <= A.goal_dis;
  if (abs(A.target_ind[b] - len) >= 5) goal = false;
  if (goal && fabs(v) <= A.stop_speed) {
    if (lane == 0) A.done[b] = 1;
    return;
  }
  // mpc_intersection.py:107: skip the update when the truncated course ends exactly at the ego's course point
  int a0 = A.agent_idx[b];
  bool update = true;
  if (A.steps[b] > 0) {
    const int last = min(len, N) - 1;
    const int at = min(a0, last);                 // tmp_trajectory[traj_agent_idx] (a0 < len always holds: cut >= a0 + 1)
    update = (cx[at] != cx[last]) || (cy[at] != cy[last]) || (cyaw[at] != cyaw[last]);
  }
  if (update) {
    const int near = nearest_index<32>(cx, cy, N, a0, x, y, lane, 0xffffffffu);
    if (near < 0) { if (lane == 0) A.done[b] = 2; return; }     // reference raises (trajectories.py:120)
    a0 = near;
  }
  if (lane == 0) A.agent_idx[b] = a0;
}

// One thread per episode: controls from the step's record, xref deviation, plant step, history, counters.
__global__ void episode_post_kernel(const EpisodeArgs A) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= A.B) return;
  double* history = A.history;
  double t_now = A.t_now;
  if (A.iter_dev) {
    const int it = *A.iter_dev;
    const bool in_rows = it < A.history_rows;
    history = (A.history_base && in_rows) ? A.history_base + (size_t)it * A.B * 8 : nullptr;
    t_now = (double)(it + 1) * A.dt_loop;
    if (A.flags_base && A.flag && in_rows) A.flags_base[(size_t)it * A.B + b] = A.flag[b];     // all episodes, as the host copy did
  }
  if (A.done[b]) return;
  const double* prm = A.params ? A.params + (size_t)b * JMPC_NPARAM : A.defaults.v;
  const double* rec = A.record + (size_t)b * JMPC_RECORD_LEN;
  const int status = (int)rec[3];
  if (status == JMPC_INDEX_RULE) { A.done[b] = 2; return; }
  const int cid = A.course_id ? A.course_id[b] : 0;
  const size_t coff = (size_t)cid * A.course_stride;
  double x = A.state[4 * b], y = A.state[4 * b + 1], v = A.state[4 * b + 2], yaw = A.state[4 * b + 3];
  const int target = (int)rec[4];
  A.target_ind[b] = target;
  double delta = A.di[b], acc = prm[JMPC_P_MAX_DECEL];
  double dev = nan("");
  // any status but OPTIMAL is a failed solve: the previous steer is kept and the ego brakes with MAX_DECEL
  // (mpc.py:298-301); that includes JMPC_MAX_ITER, whose record carries no controls
  if (status == JMPC_OPTIMAL) {
    delta = rec[0]; acc = rec[1];
    // get_current_xref_deviation (mpc.py:305-312); ox[0], oy[0] are the current position
    const double ang = A.cyaw[coff + target] + M_PI / 2;
    const double ex = A.cx[coff + target] - x, ey = A.cy[coff + target] - y;
    const double px = cos(ang) * ex, py = sin(ang) * ey;
    dev = sqrt(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)));
  }
  A.di[b] = delta;
  if (A.warm) A.warm[b] = (status == JMPC_OPTIMAL) ? 1 : 0;
  // Simulation.step (simulation.py:35-47)
  const double dt = prm[JMPC_P_DT], L = prm[JMPC_P_L], ms = prm[JMPC_P_MAX_STEER];
  const double dcl = fmax(fmin(delta, ms), -ms);
  const double xd = __dmul_rn(v, cos(yaw)), yd = __dmul_rn(v, sin(yaw)), td = __dmul_rn(v / L, tan(dcl));
  x = __dadd_rn(x, __dmul_rn(xd, dt));
  y = __dadd_rn(y, __dmul_rn(yd, dt));
  yaw = __dadd_rn(yaw, __dmul_rn(td, dt));
  v = __dadd_rn(v, __dmul_rn(acc, dt));
  v = fmax(fmin(v, prm[JMPC_P_SIM_MAX_SPEED]), prm[JMPC_P_MIN_SPEED]);
  A.state[4 * b] = x; A.state[4 * b + 1] = y; A.state[4 * b + 2] = v; A.state[4 * b + 3] = yaw;
  A.steps[b] += 1;
  if (history) {
    double* hrow = history + (size_t)b * 8;          // History.store (simulation.py:76-84), raw delta as stored there
    hrow[0] = x; hrow[1] = y; hrow[2] = yaw; hrow[3] = v; hrow[4] = t_now + dt; hrow[5] = delta; hrow[6] = acc;
    hrow[7] = dev;
  }
}

__global__ void counter_add_kernel(int* counter, int delta) { *counter += delta; }

// Constant-input obstacle motion, one thread per obstacle: obstacles [B][n_obs][6] = x, y, v, yaw, a, steer.
__global__ void obstacle_step_kernel(int count, double* __restrict__ obs, const int* __restrict__ done, int n_obs,
                                     double dt, double L) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  if (done && done[k / n_obs]) return;
  double* o = obs + (size_t)k * 6;
  double x = o[0], y = o[1], v = o[2], yaw = o[3];
  double s, c;
  sincos(yaw, &s, &c);
  x = __dadd_rn(x, __dmul_rn(__dmul_rn(v, c), dt));
  y = __dadd_rn(y, __dmul_rn(__dmul_rn(v, s), dt));
  v = __dadd_rn(v, __dmul_rn(o[4], dt));
  yaw = __dadd_rn(yaw, __dmul_rn(__dmul_rn(v / L, tan(o[5])), dt));
  o[0] = x; o[1] = y; o[2] = v; o[3] = yaw;
}

// ---- the reference's scripted obstacles (main/lib/moving_obstacles.py) -------------------------------------------
// One thread per obstacle.  `script` [count][JMPC_OBS_SCRIPT_LEN] is constant: kind, direction (+1 / -1), turning,
// speed, offset (<= 0: none), the Bicycle's sample time, the dt of the offset test (the roundabout class overwrites
// its own with 0.2, moving_obstacles.py:45, while its Bicycle keeps the constructor's), and one auxiliary value (the
// turning steer angle arctan(L / 5) of the roundabout rules, computed on the host; the arterial's initial speed).
// `model` [count][4] = Bicycle xc, yc, theta and the step counter.  `obs` [count][6] receives what `get()` returns:
// x, y, forward_velocity, theta, 0, steering_angle -- the tuple the flag kernel consumes.
//
// The scenario loop calls get() (prediction), then step() (mpc_intersection.py:123-125, 157-160).  get() evaluates
// its tuple left to right, so theta is read BEFORE the steering property runs, and the roundabout's steering property
// overwrites model.theta as a side effect (moving_obstacles.py:82-84, 98-100) -- both reproduced.
__device__ __forceinline__ double scripted_speed(const double* sc, double counter) {
  const int kind = (int)sc[JMPC_OBS_KIND];
  const double offset = sc[JMPC_OBS_OFFSET];
  const bool moving = !(offset > 0.0) || counter > __ddiv_rn(offset, sc[JMPC_OBS_DT_OFFSET]);
  if (moving) return sc[JMPC_OBS_SPEED];
  return kind == JMPC_OBS_ARTERIAL ? sc[JMPC_OBS_AUX] : 0.0;
}
__device__ __forceinline__ double scripted_steer(const double* sc, double xc, double yc, double& theta) {
  const int kind = (int)sc[JMPC_OBS_KIND];
  const bool right = sc[JMPC_OBS_DIRECTION] >= 0.0, turning = sc[JMPC_OBS_TURNING] != 0.0;
  double steer = 0.0;
  if (!turning) return steer;
  if (kind == JMPC_OBS_TINTERSECTION) {                 // moving_obstacles.py:197-213
    if (right) { if (xc >= -10.0 && theta > -M_PI / 2) steer = -0.38; }
    else if (xc <= 12.0 && theta < 3 * M_PI / 2) steer = 0.19;
  } else if (kind == JMPC_OBS_ROUNDABOUT) {             // moving_obstacles.py:66-103, the ifs in sequence
    const double ang = sc[JMPC_OBS_AUX];
    if (right) {
      if (-7.0 <= xc && xc <= -4.0 && yc < 0.0) steer = -ang;
      if (-3.0 < xc) steer = ang;
      if (yc > 0.0 && -5.0 <= xc && xc <= -3.0) steer = -ang;
      if (xc <= -3.0 && yc > 0.0) { theta = -M_PI; steer = 0.0; }
    } else {
      if (4.0 <= xc && xc <= 7.0 && yc > 0.0) steer = -ang;
      if (xc < 3.0) steer = ang;
      if (yc < 0.0 && 3.0 <= xc && xc <= 5.0) steer = -ang;
      if (3.0 <= xc && yc < 0.0) { theta = 0.0; steer = 0.0; }
    }
  }
  return steer;
}
__global__ void scripted_obstacle_kernel(int count, const double* __restrict__ script, double* __restrict__ model,
                                         double* __restrict__ obs, const int* __restrict__ done, int n_obs, int advance,
                                         double L) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  if (done && done[k / n_obs]) return;
  const double* sc = script + (size_t)k * JMPC_OBS_SCRIPT_LEN;
  double* m = model + (size_t)k * 4;
  double xc = m[0], yc = m[1], theta = m[2], counter = m[3];
  if (advance) {                                        // step(): moving_obstacles.py:113-116, bicycle/main.py:28-41
    const double steer = scripted_steer(sc, xc, yc, theta);
    const double v = scripted_speed(sc, counter), dt = sc[JMPC_OBS_DT_MODEL];
    const double xd = __dmul_rn(v, cos(theta)), yd = __dmul_rn(v, sin(theta)), td = __dmul_rn(v / L, tan(steer));
    xc = __dadd_rn(xc, __dmul_rn(xd, dt));
    yc = __dadd_rn(yc, __dmul_rn(yd, dt));
    theta = __dadd_rn(theta, __dmul_rn(td, dt));
    counter += 1.0;
  }
  double* o = obs + (size_t)k * 6;                      // get(): moving_obstacles.py:118-120
  o[0] = xc; o[1] = yc; o[2] = scripted_speed(sc, counter); o[3] = theta; o[4] = 0.0;
  o[5] = scripted_steer(sc, xc, yc, theta);             // may overwrite theta (after it has been reported)
  m[0] = xc; m[1] = yc; m[2] = theta; m[3] = counter;
}

}  // namespace jmpc
