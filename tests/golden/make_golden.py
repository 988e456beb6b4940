"""Generate the committed golden fixtures (run HERE, in the build container; not on the GPU box).

    python tests/golden/make_golden.py [--reference /root/reference]

* ghia1982_re100.json   -- the Re=100 columns (0,1) and (6,7) of the reference's data/ghia1982.dat, the only
                           golden data the reference holds for this path (bicgsimp.py:291,300).
* cavity50_oracle.json  -- the oracle (oracle/simple.py, variant bicgsimp.py exactly as committed: 50x50,
                           lid 1.0) run to its own convergence test: outer iterations, Krylov iteration totals,
                           centreline profiles, first-steps residuals.  Used to check the oracle stays what it
                           was, to compare the CUDA path on the full run, and against Ghia.
* cavity_steps_oracle.npz -- u, v, p after 3 outer iterations at 21x17 (odd, non-square) for bit-level checks.

Taichi cannot be imported here (no wheel for this Python, no network), so these come from the oracle
restatement, not from the reference itself: parity stays "pinned by the reference's own checks re-expressed"
(Ghia profiles, SciPy cross-checks), not by reference output.
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
    a = ap.parse_args()
    ghia(a.reference)
    cavity_steps()
    cavity50()
