"""Generate the committed golden fixtures (run HERE, in the build container; not on the GPU box).

    python tests/golden/make_golden.py [--reference /root/reference]

* ghia1982_re100.json   -- the Re=100 columns (0,1) and (6,7) of the reference's data/ghia1982.dat, the only
                           golden data the reference holds for this path (bicgsimp.py:291,300).
* cavity50_oracle.json  -- the oracle (oracle/simple.py, variant bicgsimp.py exactly as committed: 50x50,
                           lid 1.0) run to its own convergence test: outer iterations, Krylov iteration totals,
                           centreline profiles, first-steps residuals.  Used to check the oracle stays what it
                           was, to compare the CUDA path on the full run, and against Ghia.
* cavity_steps_oracle.npz -- u, v, p after 3 outer iterations at 21x17 (odd, non-square) for bit-level checks.
* cavity128_oracle.json -- BASELINE config 2 at its stated size: 128x128, Re = 100, run to the reference's stopping test
                           (bicgsimp.py:326) with alpha_p = 0.02 (the committed 0.1 diverges at 128^2: SURVEY.md 6).  The
                           Krylov solves go through the C port on ONE thread (oracle/c: bit-identical to the NumPy oracle
                           in serial-summation order, tests/test_oracle_cport.py, i.e. to the order the reference's loops
                           take under the serial Taichi stand-in) so that generation takes minutes, not half an hour.

These long runs come from the oracle; the oracle itself is pinned bit for bit to golden vectors produced by executing
the reference's own source files (make_reference_golden.py -> reference_*.npz, tests/test_reference_golden.py).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.simple import SimpleOracle  # noqa: E402


def ghia(reference):
    path = os.path.join(reference, 'data', 'ghia1982.dat')
    y, u = np.loadtxt(path, unpack=True, skiprows=2, usecols=(0, 1))
    x, v = np.loadtxt(path, unpack=True, skiprows=2, usecols=(6, 7))
    out = {'source': 'data/ghia1982.dat cols (0,1),(6,7); Ghia, Ghia & Shin 1982, Re=100',
           'y': y.tolist(), 'u': u.tolist(), 'x': x.tolist(), 'v': v.tolist()}
    with open(os.path.join(HERE, 'ghia1982_re100.json'), 'w') as f:
        json.dump(out, f, indent=1)


def cavity50():
    t = time.time()
    s = SimpleOracle(1.0, 1.0, 50, 50)
    s.bc['n'][1] = 1.0
    n = s.solve()
    tot = {k: int(sum(l[2][k] for l in s.log)) for k in 'uvp'}
    out = {'case': 'bicgsimp.py as committed: SIMPLESolver(1,1,50,50), bc n tangential 1.0',
           'outer_iterations': n,
           'krylov_totals': tot,
           'first_steps': [{'momentum': l[0], 'continuity': l[1], 'iters': l[2]} for l in s.log[:5]],
           'last_step': {'momentum': s.log[-1][0], 'continuity': s.log[-1][1]},
           'u_centre': s.u[26, 1:51].tolist(),      # bicgsimp.py:293
           'v_centre': s.v[1:51, 26].tolist(),      # bicgsimp.py:302
           'seconds_numpy': time.time() - t}
    with open(os.path.join(HERE, 'cavity50_oracle.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('cavity50: outer', n, tot, f'{time.time() - t:.1f}s')


def cavity128():
    from oracle import cport, krylov
    t = time.time()
    cport.load()
    cport.set_threads(1)
    s = SimpleOracle(1.0, 1.0, 128, 128)
    s.bc['n'][1] = 1.0
    s.alpha_p = 0.02

    def c_krylov(which, kind, coef, b, eps):
        fn = cport.bicgstab_solve if kind == 'bicg' else cport.cg_solve
        if kind == 'bicg' and not b.any():          # bicgsolver.py:96-98: rho == 0 at the first iteration
            s.last_iters[which], s.last_status[which] = 0, krylov.RHO_ZERO
            return np.zeros_like(b)
        x, it, st, _ = fn(coef, b, eps=eps)
        s.last_iters[which], s.last_status[which] = it, st
        return x
    s._krylov = c_krylov
    n = s.solve()
    tot = {k: int(sum(l[2][k] for l in s.log)) for k in 'uvp'}
    out = {'case': 'BASELINE config 2: SIMPLESolver(1,1,128,128), bc n tangential 1.0, alpha_p = 0.02 (others as committed in bicgsimp.py)',
           'outer_iterations': n,
           'krylov_totals': tot,
           'first_steps': [{'momentum': l[0], 'continuity': l[1], 'iters': l[2]} for l in s.log[:5]],
           'last_step': {'momentum': s.log[-1][0], 'continuity': s.log[-1][1]},
           'u_centre': s.u[65, 1:129].tolist(),      # bicgsimp.py:293 u[nx//2 + 1, 1:ny+1]
           'v_centre': s.v[1:129, 65].tolist(),      # bicgsimp.py:302
           'seconds_c_port_1_thread': time.time() - t}
    with open(os.path.join(HERE, 'cavity128_oracle.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('cavity128: outer', n, tot, f'{time.time() - t:.1f}s')


def cavity_steps():
    s = SimpleOracle(1.3, 0.9, 21, 17)
    s.bc['n'][1] = 1.0
    s.bc['w'][1] = -0.25
    s.bc['s'][1] = 0.125
    s.init()
    for _ in range(3):
        s.step()
    np.savez_compressed(os.path.join(HERE, 'cavity_steps_oracle.npz'), u=s.u, v=s.v, p=s.p, pcor=s.pcor,
                        coef_u=s.coef_u, coef_v=s.coef_v, coef_p=s.coef_p, b_u=s.b_u, b_v=s.b_v, b_p=s.b_p,
                        iters=np.array([[l[2]['u'], l[2]['v'], l[2]['p']] for l in s.log]))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--only', default='ghia,steps,cavity50,cavity128')
    a = ap.parse_args()
    only = a.only.split(',')
    if 'ghia' in only:
        ghia(a.reference)
    if 'steps' in only:
        cavity_steps()
    if 'cavity50' in only:
        cavity50()
    if 'cavity128' in only:
        cavity128()
