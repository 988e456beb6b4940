"""Channel + cylinder oracle: the stencil-form restatement (T2, oracle/karman.py) against the literal matrix
restatement of karman_taichi.py (T1, oracle/karman_literal.py), and against the survey's independent probe."""
import numpy as np
import pytest

from oracle import krylov
from oracle.karman import KarmanOracle
from oracle.karman_literal import KarmanLiteral


def stencil_from_literal(L):
    """Read T1's dict matrices back as stencil arrays in T2's layout."""
    nx, ny = L.nx, L.ny
    out = {}
    for name, A, M, b, ni, nj, stride, shape in (('u', L.Au, L.Mu, L.bu, nx + 1, ny, ny, (nx + 3, ny + 2)),
                                                 ('v', L.Av, L.Mv, L.bv, nx, ny + 1, ny + 1, (nx + 2, ny + 3)),
                                                 ('p', L.Ap, L.Mp, L.bp, nx, ny, ny, (nx + 2, ny + 2))):
        coef = np.zeros(shape + (5,)); Mv = np.zeros(shape); bb = np.zeros(shape)
        for i in range(1, ni + 1):
            for j in range(1, nj + 1):
                k = (i - 1) * stride + (j - 1)
                coef[i, j] = (A[k, k], A[k, k - stride], A[k, k + stride], A[k, k + 1], A[k, k - 1])
                Mv[i, j] = M[k, k]
                bb[i, j] = b[k]
        # nothing outside the 5-point pattern (the 7-offset product of karman_taichi.py:365 equals the 5-point stencil)
        allowed = {0, 1, -1, stride, -stride}
        stray = [(key, val) for key, val in A.d.items() if (key[1] - key[0]) not in allowed and val != 0.0]
        assert not stray, (name, stray[:3])
        out[name] = (coef, Mv, bb)
    return out


def assert_assembly_equal(L, O):
    S = stencil_from_literal(L)
    for name in 'uvp':
        coef, M, b = S[name]
        np.testing.assert_array_equal(getattr(O, f'coef_{name}'), coef, err_msg=f'coef_{name}')
        np.testing.assert_array_equal(getattr(O, f'M_{name}'), M, err_msg=f'M_{name}')
        np.testing.assert_array_equal(getattr(O, f'b_{name}'), b, err_msg=f'b_{name}')


SMALL = dict(nx=48, ny=16, lx=0.5, ly=0.1, cyl=(12, 8, 10))


def test_init_matches_literal():
    L, O = KarmanLiteral(**SMALL), KarmanOracle(**SMALL)
    L.init(); O.init()
    np.testing.assert_array_equal(L.ct, O.ct)
    np.testing.assert_array_equal(L.p, O.p)
    assert (O.ct[1:-1, 1:-1] == 1).sum() > 10          # there is a cylinder


def test_assembly_bitwise_small_with_flow():
    """Assembly of all three systems, bit for bit, on a state with non-trivial u, v, p (two sub-iterations in)."""
    L, O = KarmanLiteral(**SMALL), KarmanOracle(**SMALL)
    L.init(); O.init()
    for _ in range(2):
        O.sub_iteration(1e-8)
    for name in ('u', 'v', 'p', 'u0', 'v0', 'pcor'):
        getattr(L, name)[:] = getattr(O, name)
    # perturb so no accidental symmetry hides an index slip
    rng = np.random.default_rng(0)
    for name in ('u', 'v'):
        d = 1e-3 * rng.standard_normal(getattr(O, name).shape)
        getattr(L, name)[:] += d
        getattr(O, name)[:] += d
    L.fill_Au(); L.fill_Av(); L.fill_Ap()
    O.fill_Au(); O.fill_Av(); O.fill_Ap()
    assert_assembly_equal(L, O)
    # the documented quirk is present: some u rows carry M != diag(A)
    ic, jc = slice(1, SMALL['nx'] + 2), slice(1, SMALL['ny'] + 1)
    assert np.any(O.M_u[ic, jc] != O.coef_u[ic, jc, 0])


def test_sub_iterations_match_literal():
    L, O = KarmanLiteral(**SMALL), KarmanOracle(**SMALL)
    L.init(); O.init()
    for k in range(2):
        pl, po = L.sub_iteration(1e-8), O.sub_iteration(1e-8)
        assert po == pytest.approx(pl, rel=1e-6)
        for key in 'uvp':
            assert L.last_status[key] == O.last_status[key]
            assert abs(L.last_iters[key] - O.last_iters[key]) <= max(3, round(0.15 * L.last_iters[key])), (k, key)
    for name in ('u', 'v', 'p', 'pcor'):
        a, b = getattr(O, name), getattr(L, name)
        assert np.linalg.norm(a - b) <= 1e-6 * np.linalg.norm(b), name
    # first v solve: b = 0, x = 0 -> rho == 0 exit
    L2 = KarmanLiteral(**SMALL); L2.init(); L2.fill_Au(); L2.fill_Av()
    it, st = L2.bicgstab(L2.Av, L2.bv, L2.xv, L2.Mv, L2.ny, L2.nv, 1e-8)
    assert (it, st) == (0, 2)


def test_readme_case_320x64_assembly_and_probe_counts():
    """Config 1 (README Karman street, debug/karman_taichi.py: 320x64): T1 == T2 assembly at the initial state, and the
    first sub-iteration reproduces the survey's probe (BiCGSTAB iterations u / v / p' = 5 / rho==0 / ~555)."""
    L, O = KarmanLiteral(), KarmanOracle()
    L.init(); O.init()
    np.testing.assert_array_equal(L.ct, O.ct)
    L.fill_Au(); L.fill_Av(); L.fill_Ap()
    O.fill_Au(); O.fill_Av(); O.fill_Ap()
    assert_assembly_equal(L, O)
    O2 = KarmanOracle()
    O2.init()
    O2.sub_iteration(1e-8)
    assert O2.last_iters['u'] == 5
    assert O2.last_status['v'] == krylov.RHO_ZERO and O2.last_iters['v'] == 0
    assert abs(O2.last_iters['p'] - 555) <= 60
    assert O2.last_status['p'] == 0
