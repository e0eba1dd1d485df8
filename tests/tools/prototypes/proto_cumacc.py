"""Prototype of the condensed QP in cumulative-acceleration unknowns s_k = a_0 + ... + a_k:
the Hessian / linear term computed the way the CUDA kernel does it (oracle/condensed_model.condense_cumulative: suffix
sums of stage weights and first / second moments), checked against E' P E, E' q of the a_k formulation
(u = E [s; delta], a = D s) on instances of the bench workloads, with the conditioning of both Hessians."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
for p in (ROOT, os.path.join(ROOT, "av-simulation-at-intersections_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oracle import mpc_oracle as O                      # noqa: E402
from oracle.condensed_model import condense             # noqa: E402


from oracle.condensed_model import condense_cumulative as direct, cumulative_transform   # noqa: E402


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
            E = cumulative_transform(T)
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
