"""GPU parity with the summation ORDER pinned: the oracle adds every inner product in exactly the order the library does
(oracle/device_order.py, from kt_reduce_plan / kt_resident_plan), the element-wise arithmetic is already bit-identical
(-fmad=false), so there is no rounding left to differ: iterates, residual histories and ITERATION COUNTS must be EQUAL,
on all three solver paths (thread-block cluster, cooperative kernel, streaming kernels) and through whole SIMPLE
iterations.  This is the exact twin of every "+-k iterations" assertion elsewhere in tests/: those compare against the
oracle in NumPy's pairwise order, where BiCGSTAB's stopping iteration legitimately moves by a few counts."""
import ctypes as C

import numpy as np
import pytest

from oracle import krylov
from oracle.device_order import AutoOrder, DeviceOrder

pytestmark = pytest.mark.gpu


@pytest.fixture
def device_sums():
    yield krylov.set_summation
    krylov.set_summation('pairwise')


def sms(cuda):
    import torch
    return torch.cuda.get_device_properties(cuda).multi_processor_count


def make(cuda, coef, b, cls):
    import karman_taichi_b200 as kt
    fc = kt.Field(coef.shape, coef.dtype, cuda).from_numpy(coef)
    fb = kt.Field(b.shape, b.dtype, cuda).from_numpy(b)
    fx = kt.Field(b.shape, b.dtype, cuda)
    return cls(fc, fb, fx)


@pytest.mark.parametrize('shape', [(5, 7), (37, 53), (130, 130), (66, 323), (258, 1030), (1026, 514)])
def test_unit_inner_products_bit_exact(cuda, shape):
    """kt_dot, kt_spmv_dot, kt_bicg_s_t, kt_bicg_xr, kt_cg_xr: the sums themselves, order by order."""
    import karman_taichi_b200 as kt
    from karman_taichi_b200 import _lib
    from karman_taichi_b200.cgsolver import _p
    rng = np.random.default_rng(shape[0] + 7 * shape[1])
    coef = rng.standard_normal(shape + (5,))
    b, p, r, v, x0 = (rng.standard_normal(shape) for _ in range(5))
    s = make(cuda, coef, b, kt.BICGSolver)
    order = DeviceOrder(shape, enable=0, num_sms=sms(cuda), xr_unit=True)
    order_cg = DeviceOrder(shape, is_cg=True, enable=0, num_sms=sms(cuda))
    L, h = s._lib, s._h
    s.p.from_numpy(p); s.r.from_numpy(r); s.Ap.from_numpy(v); s.x.from_numpy(x0)
    assert s.reduce(s.p, s.r) == order('dot', p * r)
    out = (C.c_double * 2)()
    tmp = kt.Field(shape, np.float64, cuda)
    _lib.check(L.kt_spmv_dot(h, _p(s.coef), _p(s.p), _p(tmp), _p(s.b), out))
    Ap = krylov.spmv(coef, p)
    assert out[0] == order('alpha', b * Ap)
    alpha, omega = 0.43, 0.77
    _lib.check(L.kt_bicg_s_t(h, _p(s.coef), _p(s.r), _p(s.Ap), alpha, _p(s.s), _p(s.t), out))
    sv = r - alpha * v
    t = krylov.spmv(coef, sv)
    assert (out[0], out[1]) == (order('omega', t * sv), order('omega', t * t))
    _lib.check(L.kt_bicg_xr(h, _p(s.x), _p(s.r), _p(s.p), _p(s.s), _p(s.t), _p(s.b), alpha, omega, out))
    rn = sv - omega * t
    assert (out[0], out[1]) == (order('end', rn * rn), order('end', rn * b))
    s.r.from_numpy(r)
    _lib.check(L.kt_cg_xr(h, _p(s.x), _p(s.r), _p(s.p), _p(s.Ap), alpha, out))
    rc = r - alpha * v
    assert out[0] == order_cg('cg_end', rc * rc)


# grids that take each solver path with the automatic choice, plus the streaming kernels forced on the small ones
SYSTEMS = [('poisson', 14, 0.1), ('poisson', 50, 0.0), ('poisson', 128, 0.1), ('random', (60, 320), 5), ('random', (640, 128), 6),
           ('poisson', 300, 0.1), ('random', (1024, 512), 7)]


def system(kind, n, arg):
    if kind == 'poisson':
        return krylov.poisson_sin_system(n, arg)
    return krylov.random_dd_system(n[0], n[1], seed=arg)


@pytest.mark.parametrize('kind,n,arg', SYSTEMS)
@pytest.mark.parametrize('cluster', [True, False])
def test_bicgstab_solve_bit_exact(cuda, device_sums, kind, n, arg, cluster):
    import karman_taichi_b200 as kt
    coef, b = system(kind, n, arg)
    s = make(cuda, coef, b, kt.BICGSolver)
    s.cluster = cluster
    s.solve(eps=1e-9)
    device_sums('device', DeviceOrder(b.shape, enable=1 if cluster else 0, num_sms=sms(cuda)))
    x, it, st = krylov.bicgstab_solve(coef, b, eps=1e-9)
    assert (s.iterations, s.status) == (it, st)
    np.testing.assert_array_equal(s.x.to_numpy(), x)
    dev_hist = s._hist_dev[1:it + 1].cpu().numpy()
    np.testing.assert_array_equal(dev_hist, _hist_values(coef, b, it))     # sqrt(rTr) after every iteration


def _hist_values(coef, b, it):
    """sqrt(rTr) after every iteration, recomputed from the oracle (its history list holds formatted strings)."""
    vals = []
    orig = krylov.reduce

    def spy(p, q, kind='dot'):
        d = orig(p, q, kind)
        if kind == 'end' and p is q:
            vals.append(d)
        return d
    krylov.reduce = spy
    try:
        krylov.bicgstab_solve(coef, b, eps=1e-9)
    finally:
        krylov.reduce = orig
    return np.sqrt(np.array(vals[:it]))


@pytest.mark.parametrize('n,off', [(14, 0.1), (50, 0.0), (128, 0.1), (300, 0.1), (700, 0.1)])
@pytest.mark.parametrize('cluster', [True, False])
def test_cg_solve_bit_exact(cuda, device_sums, n, off, cluster):
    import karman_taichi_b200 as kt
    coef, b = krylov.poisson_sin_system(n, off)
    s = make(cuda, coef, b, kt.CGSolver)
    s.cluster = cluster
    s.solve(eps=1e-9)
    device_sums('device', DeviceOrder(b.shape, is_cg=True, enable=1 if cluster else 0, num_sms=sms(cuda)))
    hist = []
    x, it, st = krylov.cg_solve(coef, b, eps=1e-9, history=hist)
    assert (s.iterations, s.status) == (it, st)
    np.testing.assert_array_equal(s.x.to_numpy(), x)
    assert s.history == hist                         # cgsolver.py:88,111: the formatted residual history, line for line


@pytest.mark.parametrize('shape,seed', [((20, 30), 1), ((130, 130), 2), ((322, 66), 3), ((642, 130), 4), ((1026, 514), 5)])
@pytest.mark.parametrize('cluster', [True, False])
def test_jacobi_warm_start_bit_exact(cuda, device_sums, shape, seed, cluster):
    """The monolith's solver (karman_taichi.py:342-456): Jacobi plane, warm start, omega == 0 exit, monolith term order."""
    import karman_taichi_b200 as kt
    nx, ny = shape[0] - 2, shape[1] - 2
    coef, b = krylov.random_dd_system(nx, ny, seed=seed)
    rng = np.random.default_rng(seed)
    M = np.zeros(shape)
    M[1:-1, 1:-1] = coef[1:-1, 1:-1, 0] * (1.0 + 0.1 * rng.random((nx, ny)))      # M != diag(A), as in the monolith's quirk rows
    x0 = np.zeros(shape)
    x0[1:-1, 1:-1] = 0.1 * rng.standard_normal((nx, ny))
    s = make(cuda, coef, b, kt.BICGSolver)
    minv = np.zeros(shape)
    minv[1:-1, 1:-1] = 1.0 / M[1:-1, 1:-1]
    s.Minv = kt.Field(shape, np.float64, cuda).from_numpy(minv)
    s.warm_start = s.omega_zero_break = True
    s.max_steps = 100 * nx * ny
    s.cluster = cluster
    s.x.from_numpy(x0)
    s.solve(eps=1e-10)
    device_sums('device', DeviceOrder(shape, pre=True, warm=True, enable=1 if cluster else 0, num_sms=sms(cuda)))
    x, it, st = krylov.jacobi_bicgstab(coef, M, b, x0, eps=1e-10, n=nx * ny)
    assert (s.iterations, s.status) == (it, st)
    np.testing.assert_array_equal(s.x.to_numpy(), x)


@pytest.mark.parametrize('n,variant,steps', [(24, 'bicgsimp', 12), (50, 'bicgsimp', 6), (30, 'simple', 5), (200, 'bicgsimp', 2)])
def test_simple_outer_iterations_bit_exact(cuda, device_sums, n, variant, steps):
    """Whole SIMPLE outer iterations (bicgsimp.py:313-317): assembly, BCs, three Krylov solves, relaxation, corrections --
    every field equal bit for bit, every Krylov count equal."""
    import karman_taichi_b200 as kt
    from oracle.simple import SimpleOracle
    o = SimpleOracle(1.0, 1.0, n, n, variant)
    g = kt.SIMPLESolver(1.0, 1.0, n, n, variant=variant, device=cuda)
    for s in (o, g):
        s.bc['n'][1] = 1.0 if variant == 'bicgsimp' else 0.1
    device_sums('device', AutoOrder(num_sms=sms(cuda)))
    o.init(); g.init(); g._make_solvers()
    for k in range(steps):
        ro, rg = o.step(), g.step()
        assert rg == ro, k
        assert g.log[-1][2] == {key: o.last_iters[key] for key in 'uvp'}, k
    for name in ('u', 'v', 'p', 'pcor', 'mdiv', 'u_mid', 'v_mid'):
        np.testing.assert_array_equal(getattr(g, name).to_numpy(), getattr(o, name), err_msg=name)


@pytest.mark.parametrize('kw,nsub', [(dict(nx=48, ny=16, cyl=(12, 8, 10)), 4), (dict(nx=96, ny=48, cyl=(64, 32, 144)), 2),
                                     (dict(nx=320, ny=64, cyl=(64, 32, 144)), 2)])
def test_channel_sub_iterations_bit_exact(cuda, device_sums, kw, nsub):
    """Config 1 (README 320x64) and two smaller channels: whole sub-iterations of karman_taichi.py:634-653."""
    import karman_taichi_b200 as kt
    from oracle.karman import KarmanOracle
    o, g = KarmanOracle(**kw), kt.KarmanSolver(device=cuda, **kw)
    device_sums('device', AutoOrder(jacobi=True, num_sms=sms(cuda)))
    o.init(); g.init()
    for k in range(nsub):
        po, pg = o.sub_iteration(1e-8), g.sub_iteration(1e-8)
        assert pg == po, k
        assert g.last_iters == o.last_iters and g.last_status == o.last_status, k
    for name in ('u', 'v', 'p', 'pcor', 'udiv'):
        np.testing.assert_array_equal(getattr(g, name).to_numpy(), getattr(o, name), err_msg=name)


def test_channel_config4_geometry_bit_exact(cuda, device_sums):
    """BASELINE config 4's geometry at 1/8 scale: 2048 x 410 channel, cylinder scaled proportionally (centre (nx/5, ny/2),
    radius 0.1875 ny: SURVEY.md 8d), dt = 1e-5 * 320 / nx -- large enough that all three systems take the STREAMING kernels
    (Jacobi plane, warm start, the monolith's term order).  One full sub-iteration (karman_taichi.py:634-653) with the Krylov
    solves capped at 60 iterations on both sides (the p' system needs thousands at this size; bench.py caps config 4 the same
    way): assembly, every iterate and every field equal bit for bit."""
    import karman_taichi_b200 as kt
    from oracle.karman import KarmanOracle
    nx, ny = 2048, 410
    kw = dict(nx=nx, ny=ny, cyl=(nx // 5, ny // 2, int(round((0.1875 * ny) ** 2))), dt=1e-5 * 320 / nx)
    o, g = KarmanOracle(**kw), kt.KarmanSolver(device=cuda, **kw)
    o.max_iter = 60
    for s in (g.u_solver, g.v_solver, g.p_solver):
        s.max_steps = 60
    device_sums('device', AutoOrder(jacobi=True, num_sms=sms(cuda)))
    o.init(); g.init()
    np.testing.assert_array_equal(g.ct_numpy(), o.ct)
    po, pg = o.sub_iteration(1e-8), g.sub_iteration(1e-8)
    for w in 'uvp':
        np.testing.assert_array_equal(getattr(g, f'coef_{w}').to_numpy(), getattr(o, f'coef_{w}'), err_msg=f'coef_{w}')
        np.testing.assert_array_equal(getattr(g, f'b_{w}').to_numpy(), getattr(o, f'b_{w}'), err_msg=f'b_{w}')
        np.testing.assert_array_equal(getattr(g, f'M_{w}').to_numpy(), getattr(o, f'M_{w}'), err_msg=f'M_{w}')
    assert g.last_iters == o.last_iters and g.last_status == o.last_status
    assert o.last_iters['p'] == 60                                   # the cap, not convergence: the streaming kernels ran 60 bodies
    assert pg == po
    for name in ('u', 'v', 'p', 'pcor', 'udiv'):
        np.testing.assert_array_equal(getattr(g, name).to_numpy(), getattr(o, name), err_msg=name)


def test_concurrent_uv_momentum_solves_change_nothing(cuda):
    """SIMPLESolver.concurrent_uv: the u and v momentum solves on two streams at once vs one after the other
    (bicgsimp.py:205-208) -- same kernels, same deterministic sums, so every bit and every count is the same."""
    import karman_taichi_b200 as kt
    runs = []
    for conc in (True, False):
        g = kt.lid_driven_cavity(40, device=cuda)
        g.concurrent_uv = conc
        g.init(); g._make_solvers()
        for _ in range(8):
            g.step()
        runs.append(g)
    a, b = runs
    assert [l[:2] for l in a.log] == [l[:2] for l in b.log] and [l[2] for l in a.log] == [l[2] for l in b.log]
    for name in ('u', 'v', 'p', 'u_mid', 'v_mid'):
        np.testing.assert_array_equal(getattr(a, name).to_numpy(), getattr(b, name).to_numpy(), err_msg=name)
