"""Float64 numpy model of the condensed QP the CUDA kernels build and solve (TEST INFRASTRUCTURE ONLY).

The reference hands cvxpy the sparse (x, u) problem (main/lib/mpc.py:151-196).  Because the steering
operating point is always 0 (`dref == 0`, mpc.py:96) the linearised dynamics decouple into

    v_t   = v0   + dt   * sum_{k<t} a_k
    psi_t = psi0 + sum_{k<t} g_k * delta_k                       g_k = dt * vbar_k / L
    X_t   = X0 + sum_{j<t} [ alpha_j v_j - beta_j (psi_j - phibar_j) ]    alpha = dt cos(phibar), beta = dt vbar sin(phibar)
    Y_t   = Y0 + sum_{j<t} [ gamma_j v_j + kappa_j (psi_j - phibar_j) ]   gamma = dt sin(phibar), kappa = dt vbar cos(phibar)

so the states are eliminated exactly: n = 2T unknowns u = [a_0..a_{T-1}, delta_0..delta_{T-1}] and
m = 4T-1 two-sided rows (T accel boxes, T steer boxes, T-1 steer-rate rows, T speed rows as bounds on the
running sum of a).  This file exists to localise a discrepancy (condensing vs solver) and to prototype the
solver; tests compare it with the sparse oracle.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mpc_oracle import Params, stage_state_weight


@dataclass
class CondensedQP:
    P: np.ndarray       # (n, n)   1/2 u'Pu + q'u + c0
    q: np.ndarray
    c0: float
    A: np.ndarray       # (m, n) constraint rows (constant for a given T)
    lo: np.ndarray      # (m,)
    hi: np.ndarray
    S: np.ndarray       # (T+1, 4, n)  state sensitivities: x_t = x_t^free + S_t u
    xfree: np.ndarray   # (T+1, 4)


def constraint_matrix(T: int) -> np.ndarray:
    n = 2 * T
    A = np.zeros((4 * T - 1, n))
    A[:n, :n] = np.eye(n)
    for k in range(T - 1):
        A[n + k, T + k] = -1.0
        A[n + k, T + k + 1] = 1.0
    for t in range(1, T + 1):
        A[3 * T - 1 + (t - 1), :t] = 1.0          # running sum of a (speed rows, scaled by 1/dt in the bounds)
    return A


def condense(p: Params, xref: np.ndarray, xbar: np.ndarray, x0, reaches_end) -> CondensedQP:
    T, dt = p.T, p.dt
    n = 2 * T
    vb, ph = xbar[2], xbar[3]
    alpha = dt * np.cos(ph[:T])
    beta = dt * vb[:T] * np.sin(ph[:T])
    gamma = dt * np.sin(ph[:T])
    kappa = dt * vb[:T] * np.cos(ph[:T])
    g = dt * vb[:T] / p.L
    S = np.zeros((T + 1, 4, n))
    xfree = np.zeros((T + 1, 4))
    xfree[0] = x0
    for t in range(T):
        X, Y, v, psi = xfree[t]
        xfree[t + 1] = (X + alpha[t] * v - beta[t] * (psi - ph[t]), Y + gamma[t] * v + kappa[t] * (psi - ph[t]), v, psi)
        S[t + 1, 0] = S[t, 0] + alpha[t] * S[t, 2] - beta[t] * S[t, 3]
        S[t + 1, 1] = S[t, 1] + gamma[t] * S[t, 2] + kappa[t] * S[t, 3]
        S[t + 1, 2] = S[t, 2]
        S[t + 1, 3] = S[t, 3]
        S[t + 1, 2, t] += dt
        S[t + 1, 3, T + t] += g[t]
    P = np.zeros((n, n))
    q = np.zeros(n)
    c0 = 0.0
    for t in range(1, T + 1):
        Q = stage_state_weight(p, float(xref[3, t]), bool(reaches_end[t]))
        e = xfree[t] - xref[:, t]
        P += 2.0 * S[t].T @ Q @ S[t]
        q += 2.0 * S[t].T @ (Q @ e)
        c0 += float(e @ Q @ e)
    for t in range(T):
        r = p.R_end if reaches_end[t] else p.R
        P[t, t] += 2.0 * r[0]
        P[T + t, T + t] += 2.0 * r[1]
    for t in range(T - 1):
        for blk, w in ((0, p.Rd[0]), (T, p.Rd[1])):
            i, j = blk + t, blk + t + 1
            P[i, i] += 2.0 * w
            P[j, j] += 2.0 * w
            P[i, j] -= 2.0 * w
            P[j, i] -= 2.0 * w
    A = constraint_matrix(T)
    lim = p.max_dsteer * dt
    lo = np.concatenate([np.full(T, p.max_decel), np.full(T, -p.max_steer), np.full(T - 1, -lim),
                         np.full(T, (p.min_speed - x0[2]) / dt)])
    hi = np.concatenate([np.full(T, p.max_accel), np.full(T, p.max_steer), np.full(T - 1, lim),
                         np.full(T, (p.speed - x0[2]) / dt)])
    return CondensedQP(P=P, q=q, c0=c0, A=A, lo=lo, hi=hi, S=S, xfree=xfree)


def states_from_controls(c: CondensedQP, u: np.ndarray) -> np.ndarray:
    """(4, T+1) predicted states [x, y, v, yaw] of the linearised model for controls u."""
    return (c.xfree + c.S @ u).T


def objective(c: CondensedQP, u: np.ndarray) -> float:
    return float(0.5 * u @ c.P @ u + c.q @ u + c.c0)
