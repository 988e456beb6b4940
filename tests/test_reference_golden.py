"""The oracle against golden vectors produced by executing the reference's OWN source files
(tests/golden/make_reference_golden.py under the serial Taichi stand-in).

The fixtures are committed (the GPU box has no /root/reference).  The oracle in serial-summation mode (the order the
stand-in's loops take) is held to them bit for bit -- inputs, intermediate vectors, iterates and iteration counts.  This is
what pins the oracle, and through it every parity test, to the reference's own arithmetic."""
import os

import numpy as np
import pytest

from oracle import krylov
from oracle.simple import SimpleOracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def fixture(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.fail(f'{name} is missing: run tests/golden/make_reference_golden.py in a container that has /root/reference')
    return np.load(path)


@pytest.fixture
def serial_sums():
    krylov.set_summation('serial')
    yield
    krylov.set_summation('pairwise')


def test_krylov_solvers_bit_exact(serial_sums):
    g = fixture('reference_krylov.npz')
    names = sorted({k.split('/')[0] for k in g.files})
    assert names
    for name in names:
        coef, b = g[f'{name}/coef'], g[f'{name}/b']
        # first iteration, kernel by kernel (bicgsolver.py:95-121)
        r = b.copy()
        rho = krylov.reduce(r, b)
        Ap = krylov.spmv(coef, r)
        alpha = rho / krylov.reduce(b, Ap)
        s = r - alpha * Ap
        t = krylov.spmv(coef, s)
        omega = krylov.reduce(t, s) / krylov.reduce(t, t)
        x1 = alpha * r + omega * s
        r1 = s - omega * t
        for key, val in (('Ap', Ap), ('s', s), ('t', t), ('x', x1), ('r', r1)):
            np.testing.assert_array_equal(val, g[f'{name}/bicg_it1_{key}'], err_msg=f'{name} {key}')
        np.testing.assert_array_equal(np.array([rho, alpha, omega, krylov.reduce(r1, r1)]), g[f'{name}/bicg_it1_scalars'])
        # whole solves
        x, it, st = krylov.bicgstab_solve(coef, b, eps=1e-8)
        assert it == int(g[f'{name}/bicg_iters']) and st == 0
        np.testing.assert_array_equal(x, g[f'{name}/bicg_x'])
        if f'{name}/cg_x' in g.files:
            hist = []
            x, it, st = krylov.cg_solve(coef, b, eps=1e-8, history=hist)
            assert it == int(g[f'{name}/cg_iters']) and st == 0
            np.testing.assert_array_equal(x, g[f'{name}/cg_x'])
            np.testing.assert_array_equal(np.array([float(h) for h in hist]), g[f'{name}/cg_history'])


def test_cavity_outer_iterations_bit_exact(serial_sums):
    g = fixture('reference_cavity.npz')
    cases = sorted({'/'.join(k.split('/')[:2]) for k in g.files})
    assert cases
    for key in cases:
        module = key.split('/')[0]
        nx, ny, steps = (int(v) for v in g[f'{key}/nx_ny_steps'])
        o = SimpleOracle(1.0, 1.0, nx, ny, variant=module)
        for side, val in zip('wens', g[f'{key}/bc']):
            o.bc[side][0], o.bc[side][1] = float(val[0]), float(val[1])
        if key.endswith('_tight'):      # same reference methods, smaller eps arguments (make_reference_golden.py)
            o.eps_u = o.eps_v = 1e-12
            o.eps_p = 1e-13
        o.init()
        for step in range(steps):
            o.step()
            for f in ('coef_u', 'b_u', 'coef_v', 'b_v', 'coef_p', 'b_p', 'u_mid', 'v_mid', 'pcor', 'mdiv', 'u', 'v', 'p'):
                np.testing.assert_array_equal(getattr(o, f), g[f'{key}/step{step}/{f}'], err_msg=f'{key} step {step} {f}')
            if module == 'bicgsimp':
                mres, cres, nu, nv, npc = g[f'{key}/residuals'][step]
                assert (o.log[-1][0], o.log[-1][1]) == (mres, cres)
                assert (o.log[-1][2]['u'], o.log[-1][2]['v'], o.log[-1][2]['p']) == (int(nu), int(nv), int(npc))


def test_channel_against_literal_and_stencil_oracles(serial_sums):
    from oracle.karman import KarmanOracle
    from oracle.karman_literal import KarmanLiteral
    g = fixture('reference_channel.npz')
    cases = sorted({'/'.join(k.split('/')[:2]) for k in g.files})
    assert cases
    for key in cases:
        nx, ny, subs = (int(v) for v in g[f'{key}/nx_ny_subs'])
        dt = float(g[f'{key}/params'][4])
        eps = float(g[f'{key}/params'][7]) if g[f'{key}/params'].size > 7 else 1e-8
        for cls, exact in ((KarmanLiteral, True), (KarmanOracle, True)):
            o = cls(nx=nx, ny=ny, dt=dt)                       # cylinder (64, 32, 144) as hard-coded in karman_taichi.py:149
            o.init()
            np.testing.assert_array_equal(o.ct, g[f'{key}/init/ct'])
            np.testing.assert_array_equal(o.p, g[f'{key}/init/p'])
            for sub in range(subs):
                if sub == 0 and cls is KarmanOracle:
                    # the reference's n x n pointer matrices right after fill_Au / fill_Av, folded onto the five stencil slots
                    o.fill_Au(); o.fill_Av()
                    for w in 'uv':
                        np.testing.assert_array_equal(getattr(o, f'coef_{w}'), g[f'{key}/asm0/coef_{w}'], err_msg=f'{key} coef_{w}')
                        np.testing.assert_array_equal(getattr(o, f'M_{w}'), g[f'{key}/asm0/M_{w}'], err_msg=f'{key} M_{w}')
                        np.testing.assert_array_equal(getattr(o, f'b_{w}'), g[f'{key}/asm0/b_{w}'], err_msg=f'{key} b_{w}')
                pc = o.sub_iteration(eps)
                if cls is KarmanOracle:
                    for nm in ('coef_p', 'M_p', 'b_p'):
                        np.testing.assert_array_equal(getattr(o, nm), g[f'{key}/sub{sub}/{nm}'], err_msg=f'{key} sub {sub} {nm}')
                ref_pc = g[f'{key}/info'][sub][0]
                for f in ('u', 'v', 'p', 'pcor'):
                    ref = g[f'{key}/sub{sub}/{f}']
                    # literal: same matrix index arithmetic; stencil form: same five products added in the monolith's
                    # flat-index order (oracle.krylov.spmv_monolith); both: the stand-in's serial sums -> same bits
                    np.testing.assert_array_equal(getattr(o, f), ref, err_msg=f'{key} {cls.__name__} sub {sub} {f}')
                assert pc == ref_pc
                assert [o.last_iters[k] for k in 'uvp'] == [int(v) for v in g[f'{key}/info'][sub][1:4]]
        # nothing outside the five stencil offsets (the 7-offset product of karman_taichi.py:365 is a 5-point stencil)
        for which in ('asm0/extra_u', 'asm0/extra_v', 'sub0/extra_p'):
            assert g[f'{key}/{which}'].shape[0] == 0


def test_display_frames_bit_exact():
    """oracle/post.py against the display fields the reference's own Display.post_process_field (display.py:28-40) and the
    monolith's post_process_field (karman_taichi.py:544-570) produced from the final states of the fixture runs."""
    from oracle import post
    g = fixture('reference_cavity.npz')
    cases = sorted({'/'.join(k.split('/')[:2]) for k in g.files})
    for key in cases:
        steps = int(g[f'{key}/nx_ny_steps'][2])
        last = f'{key}/step{steps - 1}'
        img, panels = post.class_frame(g[f'{last}/u'], g[f'{last}/v'], g[f'{last}/p'], g[f'{last}/pcor'], g[f'{last}/mdiv'])
        for name, f in (('u', 'udisp'), ('v', 'vdisp'), ('p', 'pdisp'), ('pcor', 'pcordisp'), ('div', 'mdivdisp')):
            np.testing.assert_array_equal(panels[name], g[f'{key}/display/{f}'], err_msg=f'{key} {f}')
    c = fixture('reference_channel.npz')
    seen = 0
    for key in sorted({'/'.join(k.split('/')[:2]) for k in c.files}):
        if f'{key}/display/u_disp' not in c.files:
            continue
        seen += 1
        nx, ny, subs = (int(v) for v in c[f'{key}/nx_ny_subs'])
        lx, ly = float(c[f'{key}/params'][0]), float(c[f'{key}/params'][1])
        last = f'{key}/sub{subs - 1}'
        img, panels, udiv = post.monolith_frame(c[f'{last}/u'], c[f'{last}/v'], c[f'{last}/p'], c[f'{last}/pcor'], c[f'{last}/udiv'],
                                                lx / nx, ly / ny)
        np.testing.assert_array_equal(udiv, c[f'{key}/display/udiv'])
        for name, f in (('u', 'u_disp'), ('v', 'v_disp'), ('p', 'p_disp'), ('pcor', 'pcor_disp'), ('div', 'udiv_disp')):
            np.testing.assert_array_equal(panels[name], c[f'{key}/display/{f}'], err_msg=f'{key} {f}')
        assert img.shape == (3 * (nx + 2), 15 * (ny + 2))
    assert seen >= 1
