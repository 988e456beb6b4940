"""Run the reference's script cases on the B200 path (the reference solves at import: ``bicgsimp.py:356-371``,
``simple.py:288-303``, ``karman_taichi.py:625-683``; here it is an explicit command).

    python -m karman_taichi_b200 cavity [--n 50] [--lid 1.0] [--variant bicgsimp|simple] [--max-steps N] [--frame out.png]
    python -m karman_taichi_b200 karman [--nx 320 --ny 64] [--time-steps 1] [--max-sub 500] [--frame-prefix timestep]
    python -m karman_taichi_b200 validate pressure [--n 50] [--eps 1e-6]      # pressure_valid.py:240-281
    python -m karman_taichi_b200 validate momentum [--n 50] [--sweeps 1000]   # momentum_valid.py:210-225
"""
from __future__ import annotations

import argparse
import time


def main(argv=None):
    ap = argparse.ArgumentParser(prog='python -m karman_taichi_b200')
    sub = ap.add_subparsers(dest='case', required=True)
    c = sub.add_parser('cavity', help='lid-driven cavity, SIMPLE + BiCGSTAB/CG (bicgsimp.py / simple.py)')
    c.add_argument('--n', type=int, default=50)
    c.add_argument('--lid', type=float, default=1.0)
    c.add_argument('--variant', default='bicgsimp', choices=['bicgsimp', 'simple'])
    c.add_argument('--alpha-p', type=float, default=None)
    c.add_argument('--max-steps', type=int, default=None)
    c.add_argument('--frame', default=None, help='write the display frame (display.py) to this PNG at the end')
    c.add_argument('--dump', default=None, help='directory for the CSV dumps of simple.py:51-75')
    k = sub.add_parser('karman', help='channel + square-cell cylinder (karman_taichi.py)')
    k.add_argument('--nx', type=int, default=320)
    k.add_argument('--ny', type=int, default=64)
    k.add_argument('--time-steps', type=int, default=1)
    k.add_argument('--max-sub', type=int, default=500)
    k.add_argument('--momentum-iterations', type=int, default=1)
    k.add_argument('--frame-prefix', default=None)
    v = sub.add_parser('validate', help="the reference's validation scripts: Jacobi sweeps against the Krylov solvers, with wall-clocks")
    v.add_argument('which', choices=['pressure', 'momentum'])
    v.add_argument('--n', type=int, default=50)
    v.add_argument('--eps', type=float, default=1e-6, help='pressure: stopping tolerance of both solvers (pressure_valid.py:264,273)')
    v.add_argument('--sweeps', type=int, default=1000, help='Jacobi sweeps of the u-momentum system (momentum_valid.py:213)')
    v.add_argument('--frame-prefix', default=None, help="write the display frames '<prefix>-jacob.png' / '<prefix>-bicg.png'")
    a = ap.parse_args(argv)
    if a.case == 'validate':
        return validate(a)

    import karman_taichi_b200 as kt
    if a.case == 'cavity':
        s = kt.lid_driven_cavity(a.n, a.lid, a.variant)
        if a.alpha_p is not None:
            s.alpha_p = a.alpha_p
        s.quiet = False
        t = time.time()
        n = s.solve(max_steps=a.max_steps)
        print(f'>>> {n} outer iterations in {time.time() - t:.2f} s')
        if a.frame:
            s.disp.display(a.frame)
        if a.dump:
            s.dump_matrix(n, 'corfin', directory=a.dump)
            s.dump_coef(n, 'corfin', directory=a.dump)
    else:
        nx, ny = a.nx, a.ny
        cyl = (64, 32, 144) if (nx, ny) == (320, 64) else (nx // 5, ny // 2, int(round((0.1875 * ny) ** 2)))
        s = kt.KarmanSolver(nx=nx, ny=ny, cyl=cyl, dt=1e-5 * 320 / nx, momentum_iterations=a.momentum_iterations)
        s.quiet = False
        s.init()
        disp = kt.Display(s)
        for step in range(a.time_steps):
            t = time.time()
            n = s.time_step(a.max_sub)
            print(f'    >> [ time = {step * s.dt: .4f} ] {n} sub-iterations in {time.time() - t:.2f} s')
            if a.frame_prefix:
                disp.display(f'{a.frame_prefix}{step:06}.png')


def validate(a):
    """``pressure_valid.py`` / ``momentum_valid.py`` ``SIMPLESolver.solve``: the same system solved by point-Jacobi sweeps
    (``JacobiSolver``, ``kt_jacobi_sweep``) and by the Krylov solver, both timed, the two fields compared."""
    import numpy as np
    import torch
    import karman_taichi_b200 as kt
    s = kt.lid_driven_cavity(a.n)                      # the script tails: SIMPLESolver(1.0, 1.0, 50, 50), bc['n'][1] = 1.0
    s.init()
    s._make_solvers()
    nx, ny = s.nx, s.ny
    rel = lambda x, y: float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))
    clock = lambda: (torch.cuda.synchronize(), time.time())[1]
    if a.which == 'momentum':
        # momentum_valid.py:210-225: 1000 Jacobi sweeps (relaxed by alpha_u), then one BiCGSTAB solve (eps 1e-8) from zero
        s.compute_coef_u(); s.set_bc()
        jac = kt.JacobiSolver(s.coef_u, s.b_u, s.u, rows=(2, nx + 1), cols=(1, ny + 1), relax=s.alpha_u)
        print('Solving using Jacobian...')
        t = clock()
        jac.solve(n_iter=a.sweeps)
        ts1 = clock() - t
        u_jac = s.u.to_numpy().copy()
        print(f'>>> Iter = {jac.sweeps - 1} , eps = {jac.residual}')
        if a.frame_prefix:
            s.disp.display(f'{a.frame_prefix}-jacob.png')
        s.init()
        s.compute_coef_u(); s.set_bc()
        print('Solving using bicg...')
        t = clock()
        s.u_momentum_solver.solve(eps=1e-8, quiet=True)
        ts2 = clock() - t
        if a.frame_prefix:
            s.u.copy_from(s.u_mid)
            s.disp.display(f'{a.frame_prefix}-bicg.png')
        print(f'>>> BiCGSTAB: {s.u_momentum_solver.iterations} iterations; |u_jacobi - u_bicg| / |u_bicg| = {rel(u_jac, s.u_mid.to_numpy()):.3e}')
        print('>>> Jacobian momentum solve took', ts1, 'sec.')
        print('>>> BICG momentum solve took', ts2, 'sec.')
        return 0
    # pressure_valid.py:240-281: momentum by Jacobi sweeps, then p' by Jacobi (res < eps) and by CG (eps), both timed
    print('Solving momentum using Jacobian...')
    s.compute_coef_u(); s.set_bc()
    kt.JacobiSolver(s.coef_u, s.b_u, s.u, rows=(2, nx + 1), cols=(1, ny + 1), relax=s.alpha_u).solve(n_iter=a.sweeps)
    kt.JacobiSolver(s.coef_v, s.b_v, s.v, rows=(1, nx + 1), cols=(2, ny + 1), relax=s.alpha_u).solve(n_iter=a.sweeps)
    s.compute_coef_p()
    print('Solving pressure using Jacobian...')
    jac = kt.JacobiSolver(s.coef_p, s.b_p, s.pcor, rows=(1, nx + 1), cols=(1, ny + 1))
    t = clock()
    jac.solve(eps=a.eps, max_sweeps=10 ** 7)
    ts1 = clock() - t
    p_jac = s.pcor.to_numpy().copy()
    print('>>> Pressure correction residual =', jac.residual, f'after {jac.sweeps} sweeps')
    if a.frame_prefix:
        s.disp.display(f'{a.frame_prefix}-jacob.png')
    print('Solving pressure using BICG...')          # (the script's label; it calls CGSolver: pressure_valid.py:215-218)
    t = clock()
    s.p_correction_solver.solve(a.eps, quiet=True)
    ts2 = clock() - t
    if a.frame_prefix:
        s.disp.display(f'{a.frame_prefix}-bicg.png')
    print(f'>>> CG: {s.p_correction_solver.iterations} iterations; |p_jacobi - p_cg| / |p_cg| = {rel(p_jac, s.pcor.to_numpy()):.3e}')
    print('>>> Jacobian p correction took', ts1, 'sec.')
    print('>>> BICG p correction took', ts2, 'sec.')
    return 0


if __name__ == '__main__':
    main()
