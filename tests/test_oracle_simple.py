"""Pin the cavity SIMPLE oracle: Ghia et al. 1982 (the reference's only golden data for this path,
bicgsimp.py:291-304) and the committed oracle fixtures (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import krylov
from oracle.simple import SimpleOracle

G = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_cavity50_converged_profiles_match_ghia():
    """The full bicgsimp.py run (50x50, lid 1.0) as stored by make_golden.py against Ghia's Re=100 data.
    Thresholds: SURVEY.md section 4 probe (<= 0.0036 in u, <= 0.0084 in v on a 50x50 grid)."""
    g, c = load('ghia1982_re100.json'), load('cavity50_oracle.json')
    yc = (np.arange(50) + 0.5) / 50                       # cell-centre coordinate of u[26, 1:51] / v[1:51, 26]
    u = np.interp(g['y'], yc, c['u_centre'])
    v = np.interp(g['x'], yc, c['v_centre'])
    assert np.abs(u - np.array(g['u']))[1:-1].max() < 0.004
    assert np.abs(v - np.array(g['v']))[1:-1].max() < 0.009
    assert c['outer_iterations'] == 720                   # the survey's independent probe: converges at 719
    assert c['last_step']['momentum'] < 1e-2 and c['last_step']['continuity'] < 1e-6


def test_cavity50_first_steps_reproduce_golden():
    c = load('cavity50_oracle.json')
    s = SimpleOracle(1.0, 1.0, 50, 50)
    s.bc['n'][1] = 1.0
    s.init()
    for k in range(3):
        m, d = s.step()
        ref = c['first_steps'][k]
        assert m == pytest.approx(ref['momentum'], rel=1e-12)
        assert d == pytest.approx(ref['continuity'], rel=1e-9)
        assert s.last_iters == ref['iters']
    # the first v solve leaves through rho == 0 (b_v = 0): bicgsolver.py:96-98
    assert c['first_steps'][0]['iters']['v'] == 0


def test_steps_fixture_regression():
    z = np.load(os.path.join(G, 'cavity_steps_oracle.npz'))
    s = SimpleOracle(1.3, 0.9, 21, 17)
    s.bc['n'][1] = 1.0
    s.bc['w'][1] = -0.25
    s.bc['s'][1] = 0.125
    s.init()
    for _ in range(3):
        s.step()
    for name in ('u', 'v', 'p', 'coef_u', 'coef_v', 'coef_p', 'b_u', 'b_v', 'b_p'):
        np.testing.assert_array_equal(getattr(s, name), z[name], err_msg=name)


def test_assembled_systems_are_consistent():
    """Structural checks the reference relies on: ghost rows of coef/b stay zero (so Krylov vectors stay zero
    there), off-diagonals are non-positive at low cell Reynolds number, p' matrix rows sum to zero except the pin."""
    s = SimpleOracle(1.0, 1.0, 12, 9)
    s.bc['n'][1] = 1.0
    s.init()
    s.step()
    s.compute_coef_u(); s.compute_coef_v(); s.set_bc(); s.compute_coef_p()
    for c, b, (i0, i1, j0, j1) in ((s.coef_u, s.b_u, (2, 12, 1, 9)), (s.coef_v, s.b_v, (1, 12, 2, 9)),
                                    (s.coef_p, s.b_p, (1, 12, 1, 9))):
        mask = np.ones(b.shape, bool)
        mask[i0:i1 + 1, j0:j1 + 1] = False
        assert np.all(c[mask] == 0.0) and np.all(b[mask] == 0.0)
    cp = s.coef_p[1:13, 1:10]
    rows = cp.sum(axis=2)
    rows[0, 0] -= 1.0                                      # pinned cell (1,1): bicgsimp.py:146-151
    assert np.abs(rows).max() < 1e-12
    assert np.all(np.isfinite(s.coef_p))
    # symmetric positive semi-definite p' operator -> CG applicable (aw[i,j] == ae[i-1,j])
    np.testing.assert_allclose(s.coef_p[2:13, 2:10, 1], s.coef_p[1:12, 2:10, 2], rtol=1e-13)


def test_simple_variant_differences():
    """simple.py has no pressure pin and adds aw instead of ae on the east wall."""
    a, b = SimpleOracle(1.0, 1.0, 8, 8, 'bicgsimp'), SimpleOracle(1.0, 1.0, 8, 8, 'simple')
    for s in (a, b):
        s.dt, s.alpha_p, s.alpha_u, s.alpha_m = 1e12, 0.1, 0.8, 0.05
        s.bc['n'][1] = 1.0
        s.u[2:9, 1:9] = 0.1
        s.v[1:9, 2:9] = -0.05
        s.compute_coef_u(); s.compute_coef_v(); s.set_bc(); s.compute_coef_p()
    assert a.coef_p[1, 1, 0] == 1.0 and b.coef_p[1, 1, 0] != 1.0
    assert not np.array_equal(a.coef_v[8, 2:9, 0], b.coef_v[8, 2:9, 0])
    np.testing.assert_array_equal(a.coef_u, b.coef_u)
    assert (b.dt, b.alpha_p) == (1e12, 0.1)  # overridden above; defaults checked below
    d = SimpleOracle(1.0, 1.0, 8, 8, 'simple')
    assert (d.dt, d.alpha_p, d.alpha_u, d.alpha_m, d.eps_u, d.eps_p) == (10000, 0.001, 0.001, 0.001, 1e-1, 1e-2)
