"""Prototype of the condensed QP in cumulative-acceleration unknowns s_k = a_0 + ... + a_k:
the Hessian / linear term computed the way the CUDA kernel does it (suffix sums of stage weights and first / second
moments), checked against E' P E, E' q of oracle/condensed_model.py (u = E [s; delta], a = D s)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
for p in (ROOT, os.path.join(ROOT, "av-simulation-at-intersections_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oracle import mpc_oracle as O                      # noqa: E402
from oracle.condensed_model import condense             # noqa: E402
from oracle.mpc_oracle import stage_state_weight        # noqa: E402


def sfx(v):
    return np.cumsum(v[::-1])[::-1]


def direct(p, xref, xbar, x0, reach):
    T, dt = p.T, p.dt
    vb, ph = xbar[2], xbar[3]
    al = np.append(dt * np.cos(ph[:T]), 0.0)            # index T: 0 (s_{T-1} moves no position)
    be = dt * vb[:T] * np.sin(ph[:T])
    ga = np.append(dt * np.sin(ph[:T]), 0.0)
    ka = dt * vb[:T] * np.cos(ph[:T])
    g = dt * vb[:T] / p.L
    cb = np.concatenate([[0.0], np.cumsum(be)])          # exclusive prefix sums, t = 0..T
    ck = np.concatenate([[0.0], np.cumsum(ka)])
    cb -= cb[T // 2]; ck -= ck[T // 2]                   # centred as in the kernel
    W = np.zeros((T + 2, 4, 4))
    for t in range(1, T + 1):
        W[t] = stage_state_weight(p, float(xref[3, t]), bool(reach[t]))
    w11, w12, w22, wv, wpsi = W[:, 0, 0], W[:, 0, 1], W[:, 1, 1], W[:, 2, 2], W[:, 3, 3]
    # free response (s = 0, delta = 0)
    xf = np.zeros(T + 1); yf = np.zeros(T + 1)
    xf[0], yf[0] = x0[0], x0[1]
    for t in range(T):
        xf[t + 1] = xf[t] + al[t] * x0[2] - be[t] * (x0[3] - ph[t])
        yf[t + 1] = yf[t] + ga[t] * x0[2] + ka[t] * (x0[3] - ph[t])
    ex = np.append(xf - xref[0], 0.0); ey = np.append(yf - xref[1], 0.0)
    ev = np.append(x0[2] - xref[2], 0.0); eps = np.append(x0[3] - xref[3], 0.0)
    WeX, WeY = w11 * ex + w12 * ey, w12 * ex + w22 * ey
    # suffix sums over stages t >= m, m = 0..T+1 (index T+1: 0)
    B_, K_ = np.append(cb, 0.0), np.append(ck, 0.0)
    S11, S12, S22 = sfx(w11), sfx(w12), sfx(w22)
    M11B, M12B, M12K, M22K = sfx(w11 * B_), sfx(w12 * B_), sfx(w12 * K_), sfx(w22 * K_)
    M11BB, M12BK, M22KK = sfx(w11 * B_ * B_), sfx(w12 * B_ * K_), sfx(w22 * K_ * K_)
    SPSI, SX, SY, SE = sfx(wpsi), sfx(WeX), sfx(WeY), sfx(wpsi * eps)
    n = 2 * T
    P = np.zeros((n, n)); q = np.zeros(n)
    dt2 = dt * dt
    for i in range(T):
        for j in range(i + 1):
            m = min(i + 2, T + 1)                        # max(i, j) + 2
            P[i, j] = 2 * dt2 * (al[i + 1] * al[j + 1] * S11[m] + (al[i + 1] * ga[j + 1] + ga[i + 1] * al[j + 1]) * S12[m]
                                 + ga[i + 1] * ga[j + 1] * S22[m])
        P[i, i] += 2 * dt2 * wv[i + 1]
        m = min(i + 2, T + 1)
        q[i] = 2 * dt * (al[i + 1] * SX[m] + ga[i + 1] * SY[m]) + 2 * dt * wv[i + 1] * ev[i + 1]
    # input weights on a = D s: B = D' M D with the tridiagonal M of the a-formulation
    def Mw(i, j):
        if i >= T or j >= T or abs(i - j) > 1:
            return 0.0
        if i == j:
            r = (p.R_end if reach[i] else p.R)[0]
            nbr = (1 if (i == 0 or i == T - 1) else 2) if T >= 2 else 0
            return 2 * r + 2 * p.Rd[0] * nbr
        return -2 * p.Rd[0]
    for i in range(T):
        for j in range(max(0, i - 2), i + 1):
            P[i, j] += Mw(i, j) - Mw(i + 1, j) - Mw(i, j + 1) + Mw(i + 1, j + 1)
    for i in range(T):                                   # steer (row) x cumulative acceleration (col)
        for j in range(T):
            m = min(max(i + 1, j + 2), T + 1)
            bi, ki = cb[i + 1], ck[i + 1]
            acc = (-al[j + 1] * (M11B[m] - bi * S11[m]) - ga[j + 1] * (M12B[m] - bi * S12[m])
                   + al[j + 1] * (M12K[m] - ki * S12[m]) + ga[j + 1] * (M22K[m] - ki * S22[m]))
            P[T + i, j] = 2 * g[i] * dt * acc
    for i in range(T):                                   # steer x steer, as in the a-formulation
        for j in range(i + 1):
            m = i + 1
            bi, ki, bj, kj = cb[i + 1], ck[i + 1], cb[j + 1], ck[j + 1]
            s11 = M11BB[m] - (bi + bj) * M11B[m] + bi * bj * S11[m]
            s12a = M12BK[m] - kj * M12B[m] - bi * M12K[m] + bi * kj * S12[m]
            s12b = M12BK[m] - bj * M12K[m] - ki * M12B[m] + ki * bj * S12[m]
            s22 = M22KK[m] - (ki + kj) * M22K[m] + ki * kj * S22[m]
            P[T + i, T + j] = 2 * g[i] * g[j] * (s11 - s12a - s12b + s22 + SPSI[m])
        r = (p.R_end if reach[i] else p.R)[1]
        nbr = (1 if (i == 0 or i == T - 1) else 2) if T >= 2 else 0
        P[T + i, T + i] += 2 * r + 2 * p.Rd[1] * nbr
        if i >= 1:
            P[T + i, T + i - 1] -= 2 * p.Rd[1]
        m = i + 1
        q[T + i] = 2 * g[i] * (-(M11B[m] * 0 + sfx(WeX * B_)[m] - cb[i + 1] * SX[m]) + (sfx(WeY * K_)[m] - ck[i + 1] * SY[m]) + SE[m])
    P = np.tril(P) + np.tril(P, -1).T
    return P, q


def main():
    from junction_mpc import synth
    from helpers import default_vector, params_from_vector
    worst = 0.0
    for name, w in [("c2", synth.make_workload(2, B=64)), ("c3", synth.make_workload(3, B=64)),
                    ("T8", synth.make_sweep_sample(8, 64)), ("T25", synth.make_sweep_sample(25, 64))]:
        base = default_vector(w)
        for k in range(0, w["B"], 4):
            pv = base if w.get("params") is None else w["params"][k]
            p = params_from_vector(pv, w["T"])
            c = w["courses"][0][:int(w["course_len"][k])]
            x0 = w["state"][k]
            try:
                xref, _, reach = O.ref_trajectory(p, x0[0], x0[1], x0[2], c[:, 0], c[:, 1], c[:, 2], int(w["target_ind"][k]))
            except O.IndexRuleError:
                continue
            xbar = O.rollout(p, x0, w["oa"][k], w["od"][k])
            cq = condense(p, xref, xbar, x0, reach)
            T = p.T
            D = np.eye(T) - np.eye(T, k=-1)
            E = np.block([[D, np.zeros((T, T))], [np.zeros((T, T)), np.eye(T)]])
            Pt, qt = E.T @ cq.P @ E, E.T @ cq.q
            Pd, qd = direct(p, xref, xbar, x0, reach)
            eP = np.abs(Pd - Pt).max() / np.abs(Pt).max()
            eq = np.abs(qd - qt).max() / max(np.abs(qt).max(), 1e-30)
            worst = max(worst, eP, eq)
            if eP > 1e-9 or eq > 1e-9:
                print(name, k, "P", eP, "q", eq)
                blocks = {"ss": (slice(0, T), slice(0, T)), "ds": (slice(T, 2 * T), slice(0, T)), "dd": (slice(T, 2 * T), slice(T, 2 * T))}
                for b, (r, cc) in blocks.items():
                    print("  ", b, np.abs(Pd[r, cc] - Pt[r, cc]).max() / np.abs(Pt).max())
                print("   q s", np.abs(qd[:T] - qt[:T]).max(), "q d", np.abs(qd[T:] - qt[T:]).max())
                return
        print(name, "ok; cond P~ %.3g vs cond P %.3g" % (np.linalg.cond(Pt), np.linalg.cond(cq.P)))
    print("worst relative error", worst)


if __name__ == "__main__":
    main()
