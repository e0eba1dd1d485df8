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


# ----------------------------------------------------------------------------------------------------
# Model of the solver the CUDA kernel runs: Mehrotra predictor-corrector on the condensed QP with the
# constraint rows grouped per stage k (accel box, steer box, steer-rate row k -> k+1, speed row t = k+1).
# K = P + A' diag(w) A is assembled from that structure exactly as the kernel does it.
# ----------------------------------------------------------------------------------------------------
def _rows_apply(T, u):
    """A u, as four per-stage vectors (abox, dbox, rate, speed-sum)."""
    a, d = u[:T], u[T:]
    rate = np.zeros(T)
    rate[:T - 1] = d[1:] - d[:-1]
    return np.stack([a, d, rate, np.cumsum(a)])


def _rows_apply_T(T, t):
    """A' t for per-stage row values t[4, T] (rate row T-1 does not exist and must be 0)."""
    out = np.zeros(2 * T)
    out[:T] = t[0] + np.cumsum(t[3][::-1])[::-1]
    out[T:] = t[1]
    out[T:2 * T - 1] -= t[2][:T - 1]
    out[T + 1:] += t[2][:T - 1]
    return out


def _assemble_K(T, P, w):
    K = P.copy()
    idx = np.arange(T)
    K[idx, idx] += w[0]
    K[T + idx, T + idx] += w[1]
    for k in range(T - 1):
        K[T + k, T + k] += w[2][k]
        K[T + k + 1, T + k + 1] += w[2][k]
        K[T + k, T + k + 1] -= w[2][k]
        K[T + k + 1, T + k] -= w[2][k]
    suffix = np.cumsum(w[3][::-1])[::-1]               # suffix[i] = sum_{k >= i} w_speed[k]
    K[:T, :T] += suffix[np.maximum.outer(idx, idx)]
    return K


def ipm_solve(c: CondensedQP, max_iter: int = 40, mu_tol: float = 1e-13, s_min: float = 1e-2, lam0: float = 1.0):
    """Returns (u, iterations, converged)."""
    T = len(c.q) // 2
    n = 2 * T
    hi = np.stack([c.hi[:T], c.hi[T:2 * T], np.append(c.hi[2 * T:3 * T - 1], 1.0), c.hi[3 * T - 1:]])
    lo = np.stack([c.lo[:T], c.lo[T:2 * T], np.append(c.lo[2 * T:3 * T - 1], -1.0), c.lo[3 * T - 1:]])
    live = np.ones((4, T), bool)
    live[2, T - 1] = False                                   # there is no rate row for the last stage
    nrow = 2 * live.sum()
    u = np.zeros(n)
    z = _rows_apply(T, u)
    sh = np.maximum(hi - z, s_min)
    sl = np.maximum(z - lo, s_min)
    lh = np.where(live, lam0, 0.0)
    ll = np.where(live, lam0, 0.0)
    gscale = 1.0 + np.abs(c.q).max()
    ok = False
    it = 0
    for it in range(1, max_iter + 1):
        z = _rows_apply(T, u)
        rd = c.P @ u + c.q + _rows_apply_T(T, np.where(live, lh - ll, 0.0))
        rph = np.where(live, z + sh - hi, 0.0)
        rpl = np.where(live, -z + sl + lo, 0.0)
        mu = float((lh * sh + ll * sl)[live].sum()) / nrow
        if mu <= mu_tol and max(np.abs(rph).max(), np.abs(rpl).max()) <= 1e-9 and np.abs(rd).max() <= 1e-9 * gscale:
            ok = True
            it -= 1
            break
        w = np.where(live, lh / sh + ll / sl, 0.0)
        K = _assemble_K(T, c.P, w)
        try:
            Lc = np.linalg.cholesky(K)
        except np.linalg.LinAlgError:
            break

        def newton(rch, rcl):
            th = np.where(live, (-rch + lh * rph) / sh, 0.0)
            tl = np.where(live, (-rcl + ll * rpl) / sl, 0.0)
            rhs = -rd - _rows_apply_T(T, th - tl)
            du = np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))
            dz = _rows_apply(T, du)
            dsh = -rph - dz
            dsl = -rpl + dz
            dlh = np.where(live, (-rch - lh * dsh) / sh, 0.0)
            dll = np.where(live, (-rcl - ll * dsl) / sl, 0.0)
            return du, dsh, dsl, dlh, dll

        def max_step(v, dv):
            m = live & (dv < 0)
            return min(1.0, float((-v[m] / dv[m]).min())) if m.any() else 1.0

        du, dsh, dsl, dlh, dll = newton(lh * sh, ll * sl)
        a_aff = min(max_step(sh, dsh), max_step(sl, dsl), max_step(lh, dlh), max_step(ll, dll))
        mu_aff = float(((lh + a_aff * dlh) * (sh + a_aff * dsh) + (ll + a_aff * dll) * (sl + a_aff * dsl))[live].sum()) / nrow
        sigma = (mu_aff / mu) ** 3
        du, dsh, dsl, dlh, dll = newton(lh * sh + dsh * dlh - sigma * mu, ll * sl + dsl * dll - sigma * mu)
        def raw_step(v, dv):
            m = live & (dv < 0)
            return float((-v[m] / dv[m]).min()) if m.any() else np.inf

        frac = min(0.9999, max(0.99, 1.0 - 0.1 * (1.0 - a_aff) ** 2))   # fraction to the boundary, as in the kernel
        alpha = min(1.0, frac * min(raw_step(sh, dsh), raw_step(sl, dsl), raw_step(lh, dlh), raw_step(ll, dll)))
        u = u + alpha * du
        sh, sl = sh + alpha * dsh, sl + alpha * dsl
        lh, ll = lh + alpha * dlh, ll + alpha * dll
    return u, it, ok
