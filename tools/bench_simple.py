"""SIMPLE outer-iterations / s (the second half of BASELINE.json's metric) -- development / evidence tool.

    python tools/bench_simple.py --case cavity --n 128 --steps 20
    python tools/bench_simple.py --case karman --nx 16384 --ny 3276 --steps 3 --krylov-cap 200      # config 4

cavity: SIMPLESolver (bicgsimp.py) outer iterations of the lid-driven cavity.
karman: KarmanSolver sub-iterations (karman_taichi.py:633-652) of the channel scaled to nx x ny with the geometry of
SURVEY.md 8d config 4 (cylinder centre (nx/5, ny/2), radius 0.1875 ny, dt = 1e-5 * 320 / nx).  ``--krylov-cap`` bounds
the BiCGSTAB iterations per solve (the p' solve needs O(nx) iterations at full size); the cap and the iterations
actually spent are reported so the number can be read as outer-iters/s AND as Krylov Mcell-iters/s.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import karman_taichi_b200 as kt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default='cavity', choices=['cavity', 'karman'])
    ap.add_argument('--n', type=int, default=128)
    ap.add_argument('--nx', type=int, default=320)
    ap.add_argument('--ny', type=int, default=64)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--krylov-cap', type=int, default=0)
    ap.add_argument('--alpha-p', type=float, default=None)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    out = {'case': a.case}
    if a.case == 'cavity':
        s = kt.lid_driven_cavity(a.n, device=dev)
        if a.alpha_p is not None:
            s.alpha_p = a.alpha_p
        s.init()
        s._make_solvers()
        if a.krylov_cap:
            for sol in (s.u_momentum_solver, s.v_momentum_solver):
                sol.max_steps = a.krylov_cap
            s.p_correction_solver.steps = a.krylov_cap
        step, cells = s.step, a.n * a.n
        iters = lambda: s.log[-1][2]
        out['grid'] = [a.n, a.n]
    else:
        nx, ny = a.nx, a.ny
        cyl = (nx // 5, ny // 2, int(round((0.1875 * ny) ** 2)))
        s = kt.KarmanSolver(nx=nx, ny=ny, cyl=cyl, dt=1e-5 * 320 / nx, device=dev)
        s.init()
        if a.krylov_cap:
            for sol in (s.u_solver, s.v_solver, s.p_solver):
                sol.max_steps = a.krylov_cap
        step, cells = s.sub_iteration, nx * ny
        iters = lambda: dict(s.last_iters)
        out['grid'] = [nx, ny]
        out['cylinder'] = cyl
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    tot = {'u': 0, 'v': 0, 'p': 0}
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
        it = iters()
        for k in tot:
            tot[k] += it[k]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    kry = sum(tot.values())
    out.update({'steps': a.steps, 'seconds': dt, 'outer_iters_per_s': a.steps / dt, 'krylov_iters': tot, 'krylov_cap': a.krylov_cap,
                'krylov_mcell_iters_per_s': cells * kry / dt / 1e6, 'us_per_krylov_iter': 1e6 * dt / max(kry, 1)})
    print(json.dumps(out))


if __name__ == '__main__':
    main()
