"""Pin the Krylov oracle with the reference's own checks, re-expressed with SciPy.

* debug/ti_bicg_test.py:45-98     BiCGSTAB vs np.linalg.solve on a random 5-diagonal system, np.allclose
* benchmark/2dcg.py:129-132       np.allclose(A x, b) after CG
* benchmark/cg_to_sparse/test_cg_sparse.py:19-41  CG / BiCGSTAB residuals vs a sparse direct solve
* benchmark/cg_to_sparse/validation.py:26-45      BiCGSTAB convergence vs scipy.sparse.linalg.bicgstab
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import krylov


def test_spmv_matches_csr():
    coef, b = krylov.random_dd_system(13, 29, seed=3)
    x = np.random.default_rng(1).standard_normal(b.shape)
    A = krylov.to_csr(coef)
    np.testing.assert_allclose(krylov.spmv(coef, x).ravel(), A @ x.ravel(), rtol=1e-13, atol=1e-13)


def test_spmv_out_of_range_neighbours_are_zero():
    # bare array without a ghost ring: a neighbour outside the array contributes nothing
    rng = np.random.default_rng(0)
    coef = rng.standard_normal((5, 7, 5))
    x = rng.standard_normal((5, 7))
    y = krylov.spmv(coef, x)
    ref = np.zeros_like(x)
    for i in range(5):
        for j in range(7):
            s = coef[i, j, 0] * x[i, j]
            s += coef[i, j, 1] * (x[i - 1, j] if i > 0 else 0.0)
            s += coef[i, j, 2] * (x[i + 1, j] if i < 4 else 0.0)
            s += coef[i, j, 3] * (x[i, j + 1] if j < 6 else 0.0)
            s += coef[i, j, 4] * (x[i, j - 1] if j > 0 else 0.0)
            ref[i, j] = s
    np.testing.assert_array_equal(y, ref)


def test_bicgstab_vs_direct_random_5diag():
    """debug/ti_bicg_test.py: random diagonally dominant system, compare with a direct solve."""
    coef, b = krylov.random_dd_system(20, 16, seed=0)
    x, it, st = krylov.bicgstab_solve(coef, b, eps=1e-10)
    assert st == krylov.CONVERGED
    A = krylov.to_csr(coef).toarray()
    inner = np.zeros(b.shape, bool)
    inner[1:-1, 1:-1] = True
    k = np.flatnonzero(inner.ravel())
    xd = np.linalg.solve(A[np.ix_(k, k)], b.ravel()[k])
    assert np.allclose(x.ravel()[k], xd)                      # the reference's own criterion
    np.testing.assert_allclose(x.ravel()[k], xd, rtol=1e-7, atol=1e-9)
    assert np.all(x[~inner] == 0.0)


@pytest.mark.parametrize('n,offset', [(64, 0.0), (128, 0.1)])
def test_cg_allclose_Ax_b(n, offset):
    """benchmark/2dcg.py:132"""
    coef, b = krylov.poisson_sin_system(n, offset)
    x, it, st = krylov.cg_solve(coef, b, eps=1e-10)
    assert st == krylov.CONVERGED
    assert np.allclose(krylov.spmv(coef, x), b)


def test_cg_bicg_vs_sparse_direct():
    """benchmark/cg_to_sparse/test_cg_sparse.py (N = 128 here to keep the CPU suite fast)."""
    n = 128
    coef, b = krylov.poisson_sin_system(n, 0.1)
    A = krylov.to_csr(coef).tocsc()
    inner = np.zeros(b.shape, bool)
    inner[1:-1, 1:-1] = True
    k = np.flatnonzero(inner.ravel())
    xs = spla.spsolve(A[k][:, k], b.ravel()[k])
    xc, itc, stc = krylov.cg_solve(coef, b, eps=1e-16)
    xb, itb, stb = krylov.bicgstab_solve(coef, b, eps=1e-16)
    assert stc == krylov.CONVERGED and stb == krylov.CONVERGED
    # iteration counts recorded by the survey's probe for this family: CG needs more steps than BiCGSTAB
    assert itb < itc < 400
    np.testing.assert_allclose(xc.ravel()[k], xs, rtol=0, atol=1e-12)
    np.testing.assert_allclose(xb.ravel()[k], xs, rtol=0, atol=1e-12)


def test_known_iteration_counts():
    """Iteration counts of the S1 family (SURVEY.md section 6 probe: 140 / 96 and 141 / 82), +-1."""
    coef, b = krylov.poisson_sin_system(64, 0.0)
    _, it, _ = krylov.cg_solve(coef, b, eps=1e-6)
    assert abs(it - 140) <= 1
    _, it, _ = krylov.bicgstab_solve(coef, b, eps=1e-6)
    assert abs(it - 96) <= 1
    coef, b = krylov.poisson_sin_system(512, 0.1)
    _, it, _ = krylov.cg_solve(coef, b, eps=1e-16)
    assert abs(it - 141) <= 1
    _, it, _ = krylov.bicgstab_solve(coef, b, eps=1e-16)
    assert abs(it - 82) <= 1


def test_bicgstab_history_vs_scipy():
    """validation.py: the residual history follows SciPy's BiCGSTAB (same algorithm, x0 = 0)."""
    n = 64
    coef, b = krylov.poisson_sin_system(n, 0.0)
    hist = []
    x, it, st = krylov.bicgstab_solve(coef, b, eps=1e-6, history=hist)
    ours = np.array([float(h) for h in hist])
    A = krylov.to_csr(coef)
    res = []
    xs, info = spla.bicgstab(A, b.ravel(), rtol=1e-6, atol=0.0, x0=np.zeros(b.size),
                             callback=lambda xk: res.append(np.linalg.norm(b.ravel() - A @ xk)))
    assert info == 0
    res = np.array(res)
    # same recurrence: the early history agrees to rounding; later the two summation orders drift apart
    # (BiCGSTAB is not backward stable), so the iteration counts agree only to a few steps.
    assert abs(len(res) - it) <= 5
    np.testing.assert_allclose(ours[:40], res[:40], rtol=1e-6)
    np.testing.assert_allclose(x.ravel(), xs, atol=1e-5 * np.abs(xs).max())


def test_bicgstab_rho_zero_exit():
    """b = 0 -> rho == 0 on the first test: '>>> BICG failed at first place...' (bicgsolver.py:96-98)."""
    coef, b = krylov.poisson_sin_system(8, 0.1)
    x, it, st = krylov.bicgstab_solve(coef, np.zeros_like(b), eps=1e-8)
    assert st == krylov.RHO_ZERO and it == 0 and np.all(x == 0.0)


def test_cg_zero_rhs_raises():
    coef, b = krylov.poisson_sin_system(8, 0.1)
    with pytest.raises(ZeroDivisionError):
        krylov.cg_solve(coef, np.zeros_like(b))


def test_jacobi_bicgstab_matches_direct_and_plain():
    coef, b = krylov.random_dd_system(17, 11, seed=5)
    M = coef[:, :, 0].copy()
    x0 = np.zeros_like(b)
    x, it, st = krylov.jacobi_bicgstab(coef, M, b, x0, eps=1e-10)
    assert st == krylov.CONVERGED
    np.testing.assert_allclose(krylov.spmv(coef, x), b, atol=1e-8)
    # warm start from the solution: converges at once (residual ratio is tested on r0 = b - A x0)
    x2, it2, st2 = krylov.jacobi_bicgstab(coef, M, b, x * (1 + 1e-3), eps=1e-3)
    assert st2 == krylov.CONVERGED and it2 <= it
