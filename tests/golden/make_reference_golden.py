"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE FILES (not a restatement).

    python tests/golden/make_reference_golden.py [--reference /root/reference]

Taichi is not installable here, so ``taichi_serial_shim.py`` stands in for the ``taichi`` module (serial loops, strict
IEEE doubles; its docstring lists the semantics).  The reference files are loaded from ``/root/reference`` unmodified,
with two mechanical exceptions that do not touch any arithmetic:

* ``bicgsimp.py`` / ``simple.py`` run a whole simulation at import (``bicgsimp.py:356-371``, ``simple.py:288-303``);
  the module-level statements AFTER the class definition are not executed.  The outer loop body of
  ``SIMPLESolver.solve`` (``bicgsimp.py:314-319``, ``simple.py:270-280``) is interleaved with matplotlib / GUI calls, so
  this script calls the same methods in the same order itself.
* ``karman_taichi.py`` hard-codes ``nx = 640, ny = 128`` (``:11-12``); the two constants are replaced by the test size
  before execution (the cylinder ``(i-64)^2 + (j-32)^2 < 144`` of ``:149`` stays as written).  Its main loop lives under
  ``if __name__ == '__main__'`` (``:625-683``) and is likewise driven from here, call for call.

STATUS: run in this container (the reference is plain numerics: no subprocess / network / file writes outside
``convergence.txt``); the outputs ``reference_krylov.npz``, ``reference_cavity.npz``, ``reference_channel.npz`` are
committed, since the GPU box has no /root/reference.  ``tests/test_reference_golden.py`` holds the oracle to them bit for
bit in serial-summation mode (the order the stand-in's loops take), ``tests/test_gpu_reference_golden.py`` holds the CUDA
path to them: element-wise kernels bit for bit, solves within the north_star tolerances.

The ``*_tight`` cases repeat a case with smaller ``eps`` ARGUMENTS to the reference's own solver methods
(``BICGSolver.solve(eps=...)``, ``CGSolver.solve(eps, ...)``, ``solve_momentum_bicgstab(eps, ...)``): the reference hard-codes
``eps=1e-4`` inside ``bicg_solve_momentum_eqn`` (``bicgsimp.py:206,208``), so for those the five calls of that method's
body are issued from here, call for call.  Two runs of the same recurrence whose inner products are added in a different
order agree only to (condition number) x eps; the tight twins are what pins free-running fields to 1e-8.
"""
from __future__ import annotations

import argparse
import ast
import contextlib
import io
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import taichi_serial_shim as ti      # noqa: E402

sys.modules['taichi'] = ti


def load_reference(ref, name, keep_after_defs=False, constants=None):
    """Execute ``ref/name.py`` as module ``name``; see the module docstring for the two mechanical edits."""
    path = os.path.join(ref, name + '.py')
    tree = ast.parse(open(path).read(), filename=path)
    if not keep_after_defs:
        last_def = max(k for k, n in enumerate(tree.body) if isinstance(n, (ast.ClassDef, ast.FunctionDef)))
        tree.body = tree.body[:last_def + 1]
    if constants:
        for n in tree.body:
            if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Name) \
                    and n.targets[0].id in constants:
                n.value = ast.Constant(constants[n.targets[0].id])
        ast.fix_missing_locations(tree)
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(tree, path, 'exec'), mod.__dict__)
    return mod


def as_field(a):
    f = ti.field(ti.f64, a.shape)
    f.from_numpy(a)
    return f


def count_calls(obj, method):
    """Count calls of a bound method (driver-side instrumentation; the arithmetic is untouched)."""
    fn = getattr(obj, method)
    box = [0]

    def wrapped(*a, **k):
        box[0] += 1
        return fn(*a, **k)
    setattr(obj, method, wrapped)
    return box


# ---------------------------------------------------------------------------------------------------------
def krylov_cases(ref, out):
    from oracle import krylov
    load_reference(ref, 'cgsolver', keep_after_defs=True)
    bicg = load_reference(ref, 'bicgsolver', keep_after_defs=True)
    cg = sys.modules['cgsolver']
    systems = {
        'poisson14': krylov.poisson_sin_system(14, 0.1),
        'poisson9x0': krylov.poisson_sin_system(9, 0.0),
        'random10x13': krylov.random_dd_system(10, 13, seed=3),
        'random17x6': krylov.random_dd_system(17, 6, seed=11),
    }
    for name, (coef, b) in systems.items():
        out[f'{name}/coef'], out[f'{name}/b'] = coef, b
        # BICGSolver.solve (bicgsolver.py:88-135)
        fc, fb, fx = as_field(coef), as_field(b), as_field(np.zeros_like(b))
        s = bicg.BICGSolver(fc, fb, fx)
        n = count_calls(s, 'update_x')
        s.solve(eps=1e-8, quiet=True)
        out[f'{name}/bicg_x'] = fx.to_numpy()
        out[f'{name}/bicg_r'] = s.r.to_numpy()
        out[f'{name}/bicg_iters'] = np.int64(n[0])
        out[f'{name}/bicg_scalars'] = np.array([s.alpha[None], s.beta[None], s.omega[None], s.rho[None], s.rho_1[None]])
        # the first iteration's intermediate vectors, kernel by kernel (bicgsolver.py:95-121)
        s2 = bicg.BICGSolver(fc, fb, as_field(np.zeros_like(b)))
        s2.init()
        s2.rho[None] = s2.reduce(s2.r, s2.r_tld)
        s2.copy(s2.r, s2.p)
        s2.update_phat()
        s2.compute_Ap()
        s2.alpha[None] = s2.rho[None] / s2.reduce(s2.r_tld, s2.Ap)
        s2.update_s()
        s2.update_shat()
        s2.compute_As()
        s2.copy(s2.Ashat, s2.t)
        s2.omega[None] = s2.reduce(s2.t, s2.s) / s2.reduce(s2.t, s2.t)
        s2.update_x()
        s2.update_r()
        for v in ('Ap', 's', 't', 'x', 'r'):
            out[f'{name}/bicg_it1_{v}'] = getattr(s2, v).to_numpy()
        out[f'{name}/bicg_it1_scalars'] = np.array([s2.rho[None], s2.alpha[None], s2.omega[None], s2.reduce(s2.r, s2.r)])
        # CGSolver.solve (cgsolver.py:83-114) -- on the symmetric systems only
        if name.startswith('poisson'):
            fx = as_field(np.zeros_like(b))
            c = cg.CGSolver(fc, fb, fx)
            c.solve(eps=1e-8, quiet=True)
            out[f'{name}/cg_x'] = fx.to_numpy()
            out[f'{name}/cg_r'] = c.r.to_numpy()
            out[f'{name}/cg_iters'] = np.int64(len(c.history))     # 1 initial line + one per non-final iteration
            out[f'{name}/cg_history'] = np.array([float(h) for h in c.history])
        print(f'  krylov {name}: bicg {n[0]} its' + (f', cg {len(c.history)} its' if name.startswith('poisson') else ''))


# ---------------------------------------------------------------------------------------------------------
CAVITY_FIELDS = ('u', 'v', 'p', 'pcor', 'mdiv', 'u_mid', 'v_mid', 'coef_u', 'b_u', 'coef_v', 'b_v', 'coef_p', 'b_p')


def cavity_cases(ref, out):
    for module, cases in (('bicgsimp', [('lid10x12', 10, 12, {'n': [0.0, 1.0]}, 4),
                                         ('allbc8x7', 8, 7, {'w': [0.02, 0.3], 'e': [0.02, -0.2], 'n': [0.0, 1.0], 's': [0.0, 0.5]}, 3),
                                         ('lid10x12_tight', 10, 12, {'n': [0.0, 1.0]}, 4),
                                         ('lid21x17_tight', 21, 17, {'n': [0.0, 1.0], 'w': [0.0, -0.25], 's': [0.0, 0.125]}, 3)]),
                          ('simple', [('lid9x11', 9, 11, {'n': [0.0, 0.1]}, 3),
                                      ('allbc7x8', 7, 8, {'w': [0.01, 0.3], 'e': [0.01, -0.2], 'n': [0.0, 0.1], 's': [0.0, 0.05]}, 3)])):
        mod = load_reference(ref, module)
        for name, nx, ny, bc, steps in cases:
            s = mod.SIMPLESolver(1.0, 1.0, nx, ny)
            for k, v in bc.items():
                s.bc[k][0], s.bc[k][1] = v
            key = f'{module}/{name}'
            out[f'{key}/nx_ny_steps'] = np.array([nx, ny, steps])
            out[f'{key}/bc'] = np.array([s.bc[k] for k in 'wens'])
            s.init()
            res = []
            sink = io.StringIO()
            with contextlib.redirect_stdout(sink):
                if module == 'bicgsimp':
                    # SIMPLESolver.solve, bicgsimp.py:259-263 then the loop body :314-319
                    s.u_momentum_solver = mod.BICGSolver(s.coef_u, s.b_u, s.u_mid)
                    s.v_momentum_solver = mod.BICGSolver(s.coef_v, s.b_v, s.v_mid)
                    s.p_correction_solver = mod.CGSolver(s.coef_p, s.b_p, s.pcor)
                    nu = count_calls(s.u_momentum_solver, 'update_x')
                    nv = count_calls(s.v_momentum_solver, 'update_x')
                    npc = count_calls(s.p_correction_solver, 'update_x')
                    for step in range(steps):
                        c0 = (nu[0], nv[0], npc[0])
                        if name.endswith('_tight'):
                            # the body of bicg_solve_momentum_eqn (bicgsimp.py:199-210), eps 1e-12 instead of the literal 1e-4
                            s.compute_coef_u()
                            s.compute_coef_v()
                            s.set_bc()
                            s.u_momentum_solver.update_coef(s.coef_u, s.b_u, s.u_mid)
                            s.u_momentum_solver.solve(eps=1e-12, quiet=True)
                            s.v_momentum_solver.update_coef(s.coef_v, s.b_v, s.v_mid)
                            s.v_momentum_solver.solve(eps=1e-12, quiet=True)
                            momentum_residual = s.update_velocity()
                            s.bicg_solve_pcorrection_eqn(1e-13)
                        else:
                            momentum_residual = s.bicg_solve_momentum_eqn(1)
                            s.bicg_solve_pcorrection_eqn(1e-8)
                        s.correct_pressure()
                        s.correct_velocity()
                        continuity_residual = s.compute_mdiv()
                        res.append([momentum_residual, continuity_residual, nu[0] - c0[0], nv[0] - c0[1], npc[0] - c0[2]])
                        for f in CAVITY_FIELDS:
                            out[f'{key}/step{step}/{f}'] = getattr(s, f).to_numpy()
                else:
                    # SIMPLESolver.solve loop body, simple.py:270-280
                    for step in range(steps):
                        s.solve_momentum_eqn(1)
                        s.compute_mdiv()
                        s.solve_pcorrection_eqn()
                        s.correct_pressure()
                        s.correct_velocity()
                        s.compute_mdiv()
                        res.append([0.0, float(np.abs(s.mdiv.to_numpy()).max()), 0, 0, 0])
                        for f in CAVITY_FIELDS:
                            out[f'{key}/step{step}/{f}'] = getattr(s, f).to_numpy()
                # the display frame of the final state (display.py:28-40: Display.post_process_field of the reference's own class)
                s.disp.post_process_field()
                for f in ('udisp', 'vdisp', 'pdisp', 'pcordisp', 'mdivdisp'):
                    out[f'{key}/display/{f}'] = getattr(s.disp, f).to_numpy()
            out[f'{key}/residuals'] = np.array(res)
            print(f'  cavity {key}: {steps} outer iterations, last residuals {res[-1][:2]}, Krylov its {res[-1][2:]}')


# ---------------------------------------------------------------------------------------------------------
def stencil_of(A, M, n, stride, shape, row_of):
    """Sparse matrix of the monolith -> (coef[i,j,0..4], M[i,j]) in the (ap, aw, ae, an, as) convention, plus the
    non-zero entries that are NOT on the five stencil offsets (returned so the test can insist there are none)."""
    coef = np.zeros(shape + (5,))
    Mp = np.zeros(shape)
    offs = {0: 0, -stride: 1, stride: 2, 1: 3, -1: 4}
    extra = []
    for (r, c), val in A._d.items():
        if not (0 <= r < n):
            continue
        d = c - r
        if d in offs:
            i, j = row_of(r)
            coef[i, j, offs[d]] = val
        elif val != 0.0:
            extra.append((r, c, float(val)))
    for (r, c), val in M._d.items():
        if r == c and 0 <= r < n:
            i, j = row_of(r)
            Mp[i, j] = val
    return coef, Mp, np.array(extra, dtype=np.float64).reshape(-1, 3)


def vec_of(x, shape, row_of, n):
    a = np.zeros(shape)
    v = x.to_numpy()
    for k in range(n):
        a[row_of(k)] = v[k]
    return a


def channel_cases(ref, out):
    only = os.environ.get('KT_GOLDEN_CHANNEL_CASES')
    for name, nx, ny, subs in (('plain32x16', 32, 16, 3), ('cyl96x48', 96, 48, 2), ('plain32x16_tight', 32, 16, 3),
                               ('cyl96x48_tight', 96, 48, 1)):
        if only and name not in only.split(','):
            continue
        eps = 1e-12 if name.endswith('_tight') else 1e-8
        t0 = time.time()
        m = load_reference(ref, 'karman_taichi', keep_after_defs=True, constants={'nx': nx, 'ny': ny})
        key = f'channel/{name}'
        out[f'{key}/nx_ny_subs'] = np.array([nx, ny, subs])
        out[f'{key}/params'] = np.array([m.lx, m.ly, m.rho, m.mu, m.dt, m.velo_rel, m.p_rel, eps])
        m.init()
        out[f'{key}/init/ct'] = m.ct.to_numpy()
        out[f'{key}/init/p'] = m.p.to_numpy()
        row_u = lambda k: (k // ny + 1, k % ny + 1)
        row_v = lambda k: (k // (ny + 1) + 1, k % (ny + 1) + 1)
        row_p = lambda k: (k // ny + 1, k % ny + 1)
        n_u, n_v, n_p = (nx + 1) * ny, nx * (ny + 1), nx * ny
        info = []
        pcor_max = 0.0
        for sub in range(subs):
            sink = io.StringIO()
            with contextlib.redirect_stdout(sink):
                # main loop body, karman_taichi.py:634-653, with linalg_output = 1 so iterations can be counted
                m.fill_Au()
                m.fill_Av()
                if sub == 0:
                    cu, Mu, xu_ = stencil_of(m.Au, m.Mu, n_u, ny, (nx + 3, ny + 2), row_u)
                    cv, Mv, xv_ = stencil_of(m.Av, m.Mv, n_v, ny + 1, (nx + 2, ny + 3), row_v)
                    out[f'{key}/asm0/coef_u'], out[f'{key}/asm0/M_u'], out[f'{key}/asm0/extra_u'] = cu, Mu, xu_
                    out[f'{key}/asm0/coef_v'], out[f'{key}/asm0/M_v'], out[f'{key}/asm0/extra_v'] = cv, Mv, xv_
                    out[f'{key}/asm0/b_u'] = vec_of(m.bu, (nx + 3, ny + 2), row_u, n_u)
                    out[f'{key}/asm0/b_v'] = vec_of(m.bv, (nx + 2, ny + 3), row_v, n_v)
                m.solve_momentum_bicgstab(eps, 1, 0, 1)
                text_uv = sink.getvalue()
                m.correct_conserv()
                sink.truncate(0), sink.seek(0)
                m.solve_pcorrection_bicgstab(eps, 1)
                text_p = sink.getvalue()
                cp, Mp, xp_ = stencil_of(m.Ap, m.Mp, n_p, ny, (nx + 2, ny + 2), row_p)
                b_p = vec_of(m.bp, (nx + 2, ny + 2), row_p, n_p)
                if sub == 0:
                    pcor_max = m.puv_correction()
                    pc = pcor_max
                else:
                    pc = m.puv_correction()
            # the two momentum solves print into the same buffer: split at the restart of the iteration counter
            its = [int(l.split('Iteration')[1].split(',')[0]) for l in text_uv.splitlines() if 'Iteration' in l]
            cut = next((k for k in range(1, len(its)) if its[k] == 0), len(its))
            it_u, it_v = cut, len(its) - cut
            it_p = sum('Iteration' in l for l in text_p.splitlines())
            info.append([pc, it_u, it_v, it_p, text_uv.count('failed') + text_p.count('failed')])
            out[f'{key}/sub{sub}/coef_p'], out[f'{key}/sub{sub}/M_p'], out[f'{key}/sub{sub}/extra_p'] = cp, Mp, xp_
            out[f'{key}/sub{sub}/b_p'] = b_p
            for f in ('u', 'v', 'p', 'pcor', 'udiv'):
                out[f'{key}/sub{sub}/{f}'] = getattr(m, f).to_numpy()
            print(f'  {key} sub-iteration {sub}: pcor {pc:.6e}, Krylov its u/v/p {it_u}/{it_v}/{it_p} ({time.time() - t0:.0f} s)')
        # the monolith's frame of the final state (karman_taichi.py:544-570: post_process_field; the five *_disp canvases)
        m.post_process_field()
        for f in ('u_disp', 'v_disp', 'p_disp', 'pcor_disp', 'udiv_disp', 'udiv'):
            out[f'{key}/display/{f}'] = getattr(m, f).to_numpy()
        m.time_forward()
        out[f'{key}/after_time_forward/u0'] = m.u0.to_numpy()
        out[f'{key}/after_time_forward/v0'] = m.v0.to_numpy()
        out[f'{key}/info'] = np.array(info)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--only', default='krylov,cavity,channel')
    ap.add_argument('--merge', action='store_true', help='keep the arrays already in reference_<part>.npz')
    a = ap.parse_args()
    sys.path.insert(0, a.reference)          # bicgsolver.py does 'from cgsolver import CGSolver'
    for part, fn in (('krylov', krylov_cases), ('cavity', cavity_cases), ('channel', channel_cases)):
        if part not in a.only.split(','):
            continue
        out = {}
        print(f'{part}:')
        fn(a.reference, out)
        path = os.path.join(HERE, f'reference_{part}.npz')
        if a.merge and os.path.exists(path):
            old = np.load(path)
            out = {**{k: old[k] for k in old.files}, **out}
        np.savez_compressed(path, **out)
        print(f'  -> {path} ({os.path.getsize(path) / 1024:.0f} KiB, {len(out)} arrays)')


if __name__ == '__main__':
    main()
