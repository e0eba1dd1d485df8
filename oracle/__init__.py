"""CPU oracle for the JunctionSim MPC step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or the timed CPU baseline.  The product
(``av-simulation-at-intersections_b200/junction_mpc``) never imports this package and fails loudly when
its CUDA library is missing.

Parity status:
  * rows 2-7, 9-12 of SURVEY.md §8(a) (index rule, reference sampling, rollout, linearisation,
    collision flags, plant) are PINNED against the reference's own numpy functions, imported
    unmodified in the build container by ``tests/golden/make_golden.py``; the resulting vectors
    are committed under ``tests/golden/``.
  * row 8 (the QP solve, ``main/lib/mpc.py:141-211``) is **parity unpinned**: the arithmetic lives in
    third-party cvxpy (>=1.2.0) + ECOS (>=2.0.0), neither vendored nor installable here, and the
    reference holds no golden vectors for it.  The oracle restates the QP literally in its sparse
    (x, u) form and accepts a solution only with a float64 KKT certificate; an independent
    cross-check against scipy's bundled HiGHS QP solver runs in the CPU test-suite.
"""
