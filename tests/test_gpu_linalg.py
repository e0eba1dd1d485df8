"""The solver's warp-level building blocks (tiled symmetric matvec, Cholesky, triangular solves) against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import __graft_entry__ as g
    g.build()
    from junction_mpc import synth
    from junction_mpc.batched import BatchedMPC
    return BatchedMPC([synth.load_course("intersection")], dl=0.083, T=13, max_batch=64)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 10, 16, 26, 31, 32, 33, 40, 50, 62])
def test_tiled_factor_solve_matvec(engine, n):
    rng = np.random.default_rng(n)
    G = rng.normal(size=(n, n))
    A = G @ G.T + n * np.eye(n)
    # mimic the solver's matrices: a few huge barrier weights on the diagonal
    A[np.diag_indices(n)] += np.where(rng.random(n) < 0.3, 10.0 ** rng.uniform(3, 10, n), 0.0)
    b, x = rng.normal(size=n), rng.normal(size=n)
    sol2, prod2, ok = engine.debug_linalg(A, b, x, full=True)
    assert ok
    sol, prod = sol2[:n], prod2[:n]
    if n % 2 == 0:          # the matvec in the solver's own layout (lane k owns rows k and n/2 + k)
        np.testing.assert_allclose(prod2[n:], A @ x, rtol=1e-13, atol=1e-13 * np.abs(A).max())
    np.testing.assert_allclose(prod, A @ x, rtol=1e-13, atol=1e-13 * np.abs(A).max())
    ref = np.linalg.solve(A, b)
    np.testing.assert_allclose(sol, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())
    # backward error of the solve at working precision
    assert np.abs(A @ sol - b).max() <= 1e-11 * (np.abs(A).max() * np.abs(sol).max() + np.abs(b).max())


@pytest.mark.parametrize("n", [2, 4, 10, 16, 26, 30])
@pytest.mark.parametrize("which", [0, 1])
def test_half_warp_groups(engine, n, which):
    """Two matrices per warp, one per half (the T <= 15 kernels): each half factors and solves its own copy."""
    rng = np.random.default_rng(100 + n)
    G = rng.normal(size=(n, n))
    A = G @ G.T + n * np.eye(n)
    A[np.diag_indices(n)] += np.where(rng.random(n) < 0.3, 10.0 ** rng.uniform(3, 10, n), 0.0)
    b, x = rng.normal(size=n), rng.normal(size=n)
    sol, prod, ok = engine.debug_linalg(A, b, x, group_lanes=16, which=which)
    assert ok
    np.testing.assert_allclose(prod, A @ x, rtol=1e-13, atol=1e-13 * np.abs(A).max())
    ref = np.linalg.solve(A, b)
    np.testing.assert_allclose(sol, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())
    full, _, _ = engine.debug_linalg(A, b, x)
    assert np.array_equal(sol, full)        # the same arithmetic on 16 lanes as on 32


def test_non_positive_pivot_is_reported(engine):
    A = np.eye(8)
    A[5, 5] = -1.0
    sol, prod, ok = engine.debug_linalg(A, np.ones(8), np.ones(8))
    assert not ok
