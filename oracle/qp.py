"""Dense float64 convex-QP solver with a KKT certificate (oracle side; TEST INFRASTRUCTURE ONLY).

Stands in for ``prob.solve(solver=cvxpy.ECOS)`` at ``main/lib/mpc.py:196-197`` of the reference.
cvxpy (>=1.2.0, pyproject.toml:13) and ECOS (>=2.0.0, requirements.txt:9) are third-party, not
vendored and not installable offline, so this file restates the *published* algorithm class ECOS
belongs to -- a Mehrotra predictor-corrector primal-dual interior-point method -- for

    minimise   1/2 z'Pz + q'z + c0     subject to   A z = b,   G z <= h

followed by an active-set polish, and reports the KKT residuals of what it returns.  Because the MPC
QP is strictly convex in its free directions the optimiser is unique, so any solver whose answer passes
the certificate is a valid stand-in to far inside the parity tolerances (1e-4 abs / 1e-3 rel).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.linalg import lu_factor, lu_solve

try:
    from threadpoolctl import threadpool_limits as _limits
except Exception:          # pragma: no cover
    _limits = None


@dataclass
class QPResult:
    z: np.ndarray            # primal solution
    nu: np.ndarray           # equality multipliers
    lam: np.ndarray          # inequality multipliers (>= 0)
    obj: float               # 1/2 z'Pz + q'z + c0
    iters: int               # interior-point iterations used
    polished: bool           # True when the active-set polish produced the returned point
    kkt: dict                # residuals: stat, eq, ineq, dual, comp (all scaled, see kkt_residuals)
    ok: bool                 # kkt certificate passed at `tol`


def kkt_residuals(P, q, A, b, G, h, z, nu, lam) -> dict:
    """Scaled KKT residuals of (z, nu, lam).  Every entry is relative to the size of the terms it
    is made of, so that 1e-9 means nine correct digits regardless of the cost scaling."""
    Pz = P @ z
    Atn = A.T @ nu if A.size else np.zeros_like(z)
    Gtl = G.T @ lam if G.size else np.zeros_like(z)
    s_scale = 1.0 + max(np.abs(Pz).max(initial=0.0), np.abs(q).max(initial=0.0),
                        np.abs(Atn).max(initial=0.0), np.abs(Gtl).max(initial=0.0))
    stat = np.abs(Pz + q + Atn + Gtl).max(initial=0.0) / s_scale
    eq = 0.0
    if A.size:
        Az = A @ z
        eq = np.abs(Az - b).max(initial=0.0) / (1.0 + max(np.abs(Az).max(initial=0.0), np.abs(b).max(initial=0.0)))
    ineq = dual = comp = 0.0
    if G.size:
        Gz = G @ z
        slack = h - Gz
        p_scale = 1.0 + max(np.abs(Gz).max(initial=0.0), np.abs(h).max(initial=0.0))
        ineq = max(0.0, (-slack).max(initial=0.0)) / p_scale
        dual = max(0.0, (-lam).max(initial=0.0)) / s_scale
        # per row, either the multiplier or the slack must vanish (a product would hide a wrong active set
        # behind the cost scaling: with flat cost directions that costs digits in the controls)
        comp = np.minimum(np.abs(lam) / s_scale, np.abs(slack) / p_scale).max(initial=0.0)
    return dict(stat=float(stat), eq=float(eq), ineq=float(ineq), dual=float(dual), comp=float(comp))


def _solve_sym(K, rhs):
    """Solve a symmetric (possibly indefinite / mildly singular) system with one refinement step."""
    try:
        x = np.linalg.solve(K, rhs)
    except np.linalg.LinAlgError:
        x = np.linalg.lstsq(K, rhs, rcond=None)[0]
    r = rhs - K @ x
    try:
        x = x + np.linalg.solve(K, r)
    except np.linalg.LinAlgError:
        pass
    return x


def _eqp(P, q, A, b, G, h, act):
    """Solve the equality-constrained QP with the inequality rows `act` held at their bounds."""
    n = P.shape[0]
    act = np.array(sorted(act), dtype=int)
    Ga = G[act]
    me, ma = A.shape[0], len(act)
    K = np.zeros((n + me + ma, n + me + ma))
    K[:n, :n] = P
    K[:n, n:n + me] = A.T
    K[n:n + me, :n] = A
    K[:n, n + me:] = Ga.T
    K[n + me:, :n] = Ga
    # a tiny dual regularisation keeps the system solvable when active rows are linearly dependent
    # (e.g. a steer box and the neighbouring rate rows); refinement against the unregularised matrix
    # removes its effect on the primal solution
    Kreg = K.copy()
    idx = np.arange(n, n + me + ma)
    Kreg[idx, idx] -= 1e-10
    rhs = np.concatenate([-q, b, h[act]])
    lu = lu_factor(Kreg, check_finite=False)
    sol = lu_solve(lu, rhs, check_finite=False)
    for _ in range(3):
        sol = sol + lu_solve(lu, rhs - K @ sol, check_finite=False)
    zp = sol[:n]
    nup = sol[n:n + me]
    lamp = np.zeros(G.shape[0])
    lamp[act] = sol[n + me:]
    return zp, nup, lamp


def _polish(P, q, A, b, G, h, z, lam, s, rounds: int = 25):
    """Active-set refinement started from the interior-point guess {i : lam_i > s_i}: solve the
    equality-constrained problem, add the violated rows, drop rows whose multiplier came out negative,
    repeat until the set is stable (a primal-dual active-set iteration)."""
    act = set(np.nonzero(lam > s)[0].tolist())
    best = None
    for _ in range(rounds):
        zp, nup, lamp = _eqp(P, q, A, b, G, h, sorted(act))
        slack = h - G @ zp
        scale = 1.0 + np.abs(h).max(initial=0.0)
        lscale = 1.0 + np.abs(lamp).max(initial=0.0)
        add = set(np.nonzero(slack < -1e-12 * scale)[0].tolist()) - act
        drop = {k for k in act if lamp[k] < -1e-12 * lscale}
        best = (zp, nup, lamp)
        if not add and not drop:
            break
        act = (act | add) - drop
    return best


def solve_qp(P, q, A, b, G, h, c0: float = 0.0, tol: float = 1e-9, max_iter: int = 60) -> QPResult:
    # the matrices are ~200 x 200: multi-threaded BLAS only adds contention (10-20x slower here)
    if _limits is not None:
        with _limits(limits=1):
            return _solve_qp(P, q, A, b, G, h, c0, tol, max_iter)
    return _solve_qp(P, q, A, b, G, h, c0, tol, max_iter)


def _solve_qp(P, q, A, b, G, h, c0, tol, max_iter) -> QPResult:
    P = np.asarray(P, float)
    q = np.asarray(q, float)
    n = P.shape[0]
    A = np.asarray(A, float).reshape(-1, n)
    b = np.asarray(b, float).reshape(-1)
    G = np.asarray(G, float).reshape(-1, n)
    h = np.asarray(h, float).reshape(-1)
    me, mi = A.shape[0], G.shape[0]

    def aug_factor(W):
        K = np.zeros((n + me, n + me))
        K[:n, :n] = P + G.T @ (W[:, None] * G)
        K[:n, n:] = A.T
        K[n:, :n] = A
        return K, lu_factor(K, check_finite=False)

    def aug_solve(fac, r1, r2):
        K, lu = fac
        rhs = np.concatenate([r1, r2])
        sol = lu_solve(lu, rhs, check_finite=False)
        sol = sol + lu_solve(lu, rhs - K @ sol, check_finite=False)      # one refinement step
        return sol[:n], sol[n:]

    # --- initial point: equality-constrained least-squares start, slacks/multipliers pushed interior
    z, nu = aug_solve(aug_factor(np.ones(mi)), -q + G.T @ h, b)
    s = h - G @ z
    shift = max(0.0, -s.min(initial=0.0)) + 1.0 if (s.min(initial=1.0) <= 1e-8) else 0.0
    s = s + shift
    lam = np.ones(mi)

    it = 0
    for it in range(1, max_iter + 1):
        r_d = P @ z + q + A.T @ nu + G.T @ lam
        r_e = A @ z - b
        r_p = G @ z + s - h
        mu = float(lam @ s) / max(mi, 1)
        scale_d = 1.0 + max(np.abs(q).max(initial=0.0), np.abs(P @ z).max(initial=0.0))
        scale_p = 1.0 + np.abs(h).max(initial=0.0)
        if (np.abs(r_d).max(initial=0.0) <= 1e-11 * scale_d and np.abs(r_e).max(initial=0.0) <= 1e-11 * scale_p
                and np.abs(r_p).max(initial=0.0) <= 1e-11 * scale_p and mu <= 1e-13):
            break
        # Complementarity far below its target with the primal rows satisfied: the interior-point phase has done
        # its job (identify the active set).  What is left in r_d is multiplier noise on rows whose slack is at
        # roundoff; iterating on only amplifies it (w = lam / s overflows the factorisation).  The active-set
        # refinement below produces the certified point.
        if mu <= 1e-15 and np.abs(r_e).max(initial=0.0) <= 1e-9 * scale_p and np.abs(r_p).max(initial=0.0) <= 1e-9 * scale_p:
            break
        fac = aug_factor(lam / s)

        def newton(r_c):
            t = (-r_c + lam * r_p) / s
            dz, dnu = aug_solve(fac, -r_d - G.T @ t, -r_e)
            ds = -r_p - G @ dz
            dlam = (-r_c - lam * ds) / s
            return dz, dnu, ds, dlam

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float((-v[neg] / dv[neg]).min(initial=np.inf)))

        # predictor
        dz_a, dnu_a, ds_a, dlam_a = newton(lam * s)
        alpha_a = min(max_step(s, ds_a), max_step(lam, dlam_a))
        mu_a = float((lam + alpha_a * dlam_a) @ (s + alpha_a * ds_a)) / max(mi, 1)
        sigma = (mu_a / mu) ** 3 if mu > 0 else 0.0
        # corrector
        dz, dnu, ds, dlam = newton(lam * s + ds_a * dlam_a - sigma * mu)
        alpha_p = min(1.0, 0.995 * max_step(s, ds) if np.any(ds < 0) else 1.0)
        alpha_d = min(1.0, 0.995 * max_step(lam, dlam) if np.any(dlam < 0) else 1.0)
        alpha = min(alpha_p, alpha_d)
        z = z + alpha * dz
        nu = nu + alpha * dnu
        s = s + alpha * ds
        lam = lam + alpha * dlam

    best = (z, nu, lam, False)
    res = kkt_residuals(P, q, A, b, G, h, z, nu, lam)
    try:
        zp, nup, lamp = _polish(P, q, A, b, G, h, z, lam, s)
        resp = kkt_residuals(P, q, A, b, G, h, zp, nup, lamp)
        if max(resp.values()) < max(res.values()):
            best, res = (zp, nup, lamp, True), resp
    except np.linalg.LinAlgError:
        pass
    z, nu, lam, polished = best
    obj = float(0.5 * z @ P @ z + q @ z + c0)
    return QPResult(z=z, nu=nu, lam=lam, obj=obj, iters=it, polished=polished, kkt=res,
                    ok=max(res.values()) <= tol)
