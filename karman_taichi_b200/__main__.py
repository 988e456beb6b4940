"""Run the reference's script cases on the B200 path (the reference solves at import: ``bicgsimp.py:356-371``,
``simple.py:288-303``, ``karman_taichi.py:625-683``; here it is an explicit command).

    python -m karman_taichi_b200 cavity [--n 50] [--lid 1.0] [--variant bicgsimp|simple] [--max-steps N] [--frame out.png]
    python -m karman_taichi_b200 karman [--nx 320 --ny 64] [--time-steps 1] [--max-sub 500] [--frame-prefix timestep]
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
    a = ap.parse_args(argv)

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


if __name__ == '__main__':
    main()
