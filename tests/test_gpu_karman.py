"""GPU parity: channel + cylinder (karman_taichi.py) kernels and driver vs the stencil-form oracle (T2), which is itself
proven bit-equal to the literal matrix oracle (tests/test_oracle_karman.py)."""
import numpy as np
import pytest

from oracle.karman import KarmanOracle

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


CASES = [dict(nx=48, ny=16, cyl=(12, 8, 10)), dict(nx=80, ny=33, cyl=(20, 16, 30)), dict(nx=320, ny=64, cyl=(64, 32, 144))]


def pair(cuda, kw):
    import karman_taichi_b200 as kt
    o = KarmanOracle(**kw)
    g = kt.KarmanSolver(device=cuda, **kw)
    o.init(); g.init()
    return o, g


def sync_state(o, g, rng=None):
    for name in ('u', 'v', 'p', 'u0', 'v0', 'pcor'):
        a = getattr(o, name)
        if rng is not None and name in ('u', 'v', 'u0', 'v0'):
            a += 1e-2 * rng.standard_normal(a.shape)
        getattr(g, name).from_numpy(a)
    g.x_u.from_numpy(o.x_u); g.x_v.from_numpy(o.x_v)


@pytest.mark.parametrize('kw', CASES)
def test_init_bit_exact(cuda, kw):
    o, g = pair(cuda, kw)
    np.testing.assert_array_equal(g.ct_numpy(), o.ct)
    for name in ('p', 'u', 'v', 'u0', 'v0'):
        np.testing.assert_array_equal(getattr(g, name).to_numpy(), getattr(o, name), err_msg=name)


@pytest.mark.parametrize('kw', CASES)
def test_assembly_bit_exact(cuda, kw):
    o, g = pair(cuda, kw)
    o.sub_iteration(1e-8)
    sync_state(o, g, np.random.default_rng(1))
    o.fill_Au(); o.fill_Av(); o.fill_Ap()
    g.fill_Au(); g.fill_Ap()
    for s in 'uvp':
        np.testing.assert_array_equal(getattr(g, f'coef_{s}').to_numpy(), getattr(o, f'coef_{s}'), err_msg=f'coef_{s}')
        np.testing.assert_array_equal(getattr(g, f'b_{s}').to_numpy(), getattr(o, f'b_{s}'), err_msg=f'b_{s}')
        np.testing.assert_array_equal(getattr(g, f'M_{s}').to_numpy(), getattr(o, f'M_{s}'), err_msg=f'M_{s}')
        M = getattr(o, f'M_{s}')
        with np.errstate(divide='ignore'):
            minv = np.where(M != 0.0, 1.0 / np.where(M != 0.0, M, 1.0), 0.0)
        np.testing.assert_array_equal(getattr(g, f'Minv_{s}').to_numpy(), minv, err_msg=f'Minv_{s}')


@pytest.mark.parametrize('kw', CASES[:2])
def test_relax_conserv_correct(cuda, kw):
    o, g = pair(cuda, kw)
    o.sub_iteration(1e-8)
    rng = np.random.default_rng(2)
    o.x_u[1:-1, 1:-1] += 1e-2 * rng.standard_normal(o.x_u[1:-1, 1:-1].shape)
    sync_state(o, g, rng)
    nx, ny = o.nx, o.ny
    # xu_back / xv_back: bit-exact
    U = o.u[1:nx + 2, 1:ny + 1]; U[:] = U + o.velo_rel * (o.x_u[1:nx + 2, 1:ny + 1] - U)
    V = o.v[1:nx + 1, 1:ny + 2]; V[:] = V + o.velo_rel * (o.x_v[1:nx + 1, 1:ny + 2] - V)
    g.xu_back()
    np.testing.assert_array_equal(g.u.to_numpy(), o.u)
    np.testing.assert_array_equal(g.v.to_numpy(), o.v)
    # correct_conserv: the two column sums are order-dependent -> 1e-14
    o.correct_conserv()
    g.correct_conserv()
    np.testing.assert_allclose(g.u.to_numpy(), o.u, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(g.udiv.to_numpy(), o.udiv, rtol=0, atol=1e-13 * np.abs(o.u).max())
    # puv_correction: bit-exact given identical inputs
    sync_state(o, g)
    o.fill_Au(); o.fill_Av(); g.fill_Au()
    po, pg = o.puv_correction(), g.puv_correction()
    assert pg == po
    for name in ('u', 'v', 'p'):
        np.testing.assert_array_equal(getattr(g, name).to_numpy(), getattr(o, name), err_msg=name)
    o.time_forward(); g.time_forward()
    np.testing.assert_array_equal(g.u0.to_numpy(), o.u0)


@pytest.mark.parametrize('kw,nsub', [(CASES[0], 4), (CASES[2], 3)])
def test_sub_iterations_vs_oracle(cuda, kw, nsub):
    """Config 1 (README case 320x64) and a small channel: first sub-iterations of the first time step, eps = 1e-8."""
    o, g = pair(cuda, kw)
    for k in range(nsub):
        po, pg = o.sub_iteration(1e-8), g.sub_iteration(1e-8)
        assert pg == pytest.approx(po, rel=1e-6)
        for key in 'uvp':
            assert g.last_status[key] == o.last_status[key], (k, key)
            # NumPy's pairwise order; with the order pinned the counts are EQUAL: tests/test_gpu_exact.py::test_channel_sub_iterations_bit_exact
            assert abs(g.last_iters[key] - o.last_iters[key]) <= max(3, round(0.15 * o.last_iters[key])), (k, key)
    errs = {name: rel_l2(getattr(g, name).to_numpy(), getattr(o, name)) for name in ('u', 'v', 'p', 'pcor')}
    assert errs['u'] < 1e-8 and errs['p'] < 1e-8 and errs['v'] < 1e-7 and errs['pcor'] < 1e-6, errs


def test_time_step_small_channel(cuda):
    """A whole time step (inner loop to the reference's pcor criterion, time_forward) on the small channel."""
    o, g = pair(cuda, CASES[0])
    no, ng = o.time_step(max_sub=40), g.time_step(max_sub=40)
    assert abs(no - ng) <= 1
    np.testing.assert_array_equal(g.u0.to_numpy(), g.u.to_numpy())
    if no == ng:
        assert rel_l2(g.u.to_numpy(), o.u) < 1e-7
