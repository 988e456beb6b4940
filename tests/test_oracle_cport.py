"""The C/OpenMP port (CPU baseline of bench.py) must be the same algorithm as the NumPy oracle."""
import os

import numpy as np
import pytest

from oracle import cport, krylov


def test_spmv_bit_exact_and_reduce():
    rng = np.random.default_rng(0)
    coef = rng.standard_normal((37, 53, 5))
    x = rng.standard_normal((37, 53))
    np.testing.assert_array_equal(cport.spmv(coef, x), krylov.spmv(coef, x))
    y = rng.standard_normal((37, 53))
    assert cport.reduce(x, y) == pytest.approx(float(np.sum(x * y)), rel=1e-13)


@pytest.mark.parametrize('n,offset,eps', [(64, 0.0, 1e-6), (128, 0.1, 1e-10)])
def test_solvers_match_numpy_oracle(n, offset, eps):
    coef, b = krylov.poisson_sin_system(n, offset)
    for c_fn, o_fn in ((cport.cg_solve, krylov.cg_solve), (cport.bicgstab_solve, krylov.bicgstab_solve)):
        xo, ito, sto = o_fn(coef, b, eps=eps)
        xc, itc, stc, rel = c_fn(coef, b, eps=eps)
        assert stc == sto == 0
        # CG: +-1.  BiCGSTAB: the OpenMP reduction order (like Taichi's) changes the stopping iteration by a few
        # steps from run to run -- see test_bicgstab_count_noise_floor
        assert abs(itc - ito) <= (1 if c_fn is cport.cg_solve else max(4, round(0.15 * ito)))
        assert np.linalg.norm(xc - xo) / np.linalg.norm(xo) < max(1e-8, 10 * eps)
        assert rel < eps


def test_status_codes():
    coef, b = krylov.poisson_sin_system(8, 0.1)
    x, it, st, _ = cport.bicgstab_solve(coef, np.zeros_like(b))
    assert st == krylov.RHO_ZERO and it == 0
    x, it, st, _ = cport.bicgstab_solve(coef, b, eps=1e-30, max_iter=5)
    assert st == krylov.MAX_ITER and it == 5


def test_bicgstab_count_noise_floor():
    """The stopping iteration of BiCGSTAB depends on the summation ORDER of its inner products: the NumPy oracle
    (pairwise np.sum) and the C port (sequential / per-thread sums) run the same recurrence on the same data and
    still stop a few iterations apart on some systems.  This is the noise floor the GPU parity tests allow for;
    the reference's own `sum +=` reduction is thread-order dependent, so it has the same spread against itself."""
    spread = []
    for n, off, eps in [(64, 0.0, 1e-6), (128, 0.0, 1e-10), (256, 0.0, 1e-10), (512, 0.1, 1e-12)]:
        coef, b = krylov.poisson_sin_system(n, off)
        _, it_np, st = krylov.bicgstab_solve(coef, b, eps=eps)
        _, it_c, st_c, _ = cport.bicgstab_solve(coef, b, eps=eps)
        assert st == st_c == 0
        assert abs(it_np - it_c) <= max(4, round(0.15 * it_np))
        spread.append(abs(it_np - it_c))
        # CG is stable: +-1
        _, ic_np, _ = krylov.cg_solve(coef, b, eps=eps)
        _, ic_c, _, _ = cport.cg_solve(coef, b, eps=eps)
        assert abs(ic_np - ic_c) <= 1
    print('bicgstab iteration spread numpy-vs-C:', spread)


def test_serial_summation_is_bit_identical_to_one_thread_port():
    """Two independent restatements of the same recurrence (NumPy arrays vs C loops) with the same summation order
    -- strictly sequential, the order a serial run of the reference's `sum +=` loop takes -- must agree bit for bit,
    iteration counts included: everything else in them (SpMV order, update order, host scalar algebra) is then equal."""
    krylov.set_summation('serial')
    cport.set_threads(1)
    try:
        for (coef, b), eps in [(krylov.poisson_sin_system(48, 0.1), 1e-10), (krylov.random_dd_system(37, 53, seed=2), 1e-9),
                               (krylov.poisson_sin_system(33, 0.0), 1e-8)]:
            xo, ito, sto = krylov.bicgstab_solve(coef, b, eps=eps)
            xc, itc, stc, _ = cport.bicgstab_solve(coef, b, eps=eps)
            assert (ito, sto) == (itc, stc)
            np.testing.assert_array_equal(xc, xo)
        coef, b = krylov.poisson_sin_system(40, 0.1)
        xo, ito, sto = krylov.cg_solve(coef, b, eps=1e-10)
        xc, itc, stc, _ = cport.cg_solve(coef, b, eps=1e-10)
        assert (ito, sto) == (itc, stc)
        np.testing.assert_array_equal(xc, xo)
    finally:
        krylov.set_summation('pairwise')
        cport.set_threads(os.cpu_count() or 1)
