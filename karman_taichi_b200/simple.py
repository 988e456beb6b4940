"""``SIMPLESolver`` -- host mirror of the reference class (``bicgsimp.py:10-353`` / ``simple.py:10-285``).

Same constructor ``SIMPLESolver(lx, ly, nx, ny)``, same attributes (``bc`` dict, ``alpha_p/u/m``, ``dt``, ``rho``,
``mu``, fields ``u, v, p, pcor, mdiv, coef_*, b_*``) and the same methods; each former ``@ti.kernel`` is one C-ABI
call into libkarman_b200 and the Krylov solves are device-resident.  ``variant='bicgsimp'`` (default) follows
``bicgsimp.py``; ``variant='simple'`` follows ``simple.py`` (parameters, missing pressure pin, east-wall
coefficient slip, fixed 2000 steps: SURVEY.md A.3).  Importing this module runs nothing -- the reference
solves at import (``bicgsimp.py:356-371``); use ``lid_driven_cavity()`` for that script's set-up.

GUI / matplotlib live plotting (``bicgsimp.py:268-353``) is out of scope; ``solve()`` keeps the loop, the
residual print-out and the convergence test.
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib
from .bicgsolver import BICGSolver
from .cgsolver import CGSolver, _Handle, _p
from .display import Display
from .field import Field


class SIMPLESolver:
    def __init__(self, lx, ly, nx, ny, variant='bicgsimp', dtype=np.float64, device=None):
        if variant not in ('bicgsimp', 'simple'):
            raise ValueError("variant must be 'bicgsimp' or 'simple'")
        self.variant = variant
        self.lx = lx
        self.ly = ly
        self.nx = nx
        self.ny = ny
        self.dx = self.lx / self.nx
        self.dy = self.ly / self.ny
        self.rho = 1.00
        self.mu = 0.01
        self.real = np.dtype(dtype)
        if variant == 'bicgsimp':          # bicgsimp.py:20-25
            self.dt = 1e12
            self.alpha_p, self.alpha_u, self.alpha_m = 0.1, 0.8, 0.05
            self.eps_u = self.eps_v = 1e-4   # bicgsimp.py:206,208
            self.eps_p = 1e-8                # bicgsimp.py:314
        else:                              # simple.py:20-25
            self.dt = 10000
            self.alpha_p, self.alpha_u, self.alpha_m = 0.001, 0.001, 0.001
            self.eps_u = self.eps_v = 1e-1   # simple.py:248-249
            self.eps_p = 1e-2                # simple.py:255
        mk = lambda *s: Field(s, self.real, device)
        self.u, self.v = mk(nx + 3, ny + 2), mk(nx + 2, ny + 3)
        self.u_mid, self.v_mid = mk(nx + 3, ny + 2), mk(nx + 2, ny + 3)
        self.u0, self.v0 = mk(nx + 3, ny + 2), mk(nx + 2, ny + 3)
        self.p, self.pcor, self.mdiv = mk(nx + 2, ny + 2), mk(nx + 2, ny + 2), mk(nx + 2, ny + 2)
        self.bc = {'w': [0.0, 0.0], 'e': [0.0, 0.0], 'n': [0.0, 0.0], 's': [0.0, 0.0]}
        self.ct = None                      # allocated but unused in the reference (bicgsimp.py:40)
        self.coef_u, self.b_u = mk(nx + 3, ny + 2, 5), mk(nx + 3, ny + 2)
        self.coef_v, self.b_v = mk(nx + 2, ny + 3, 5), mk(nx + 2, ny + 3)
        self.coef_p, self.b_p = mk(nx + 2, ny + 2, 5), mk(nx + 2, ny + 2)
        self._hd = _Handle(self.p)
        self._lib = self._hd.lib
        self.u_momentum_solver = None
        self.v_momentum_solver = None
        self.p_correction_solver = None
        self.disp = Display(self)           # simple.py:49
        self.quiet = True
        # The u- and v-momentum systems are independent once assembled (bicgsimp.py:202-208 solves them one after the other
        # only because Python is sequential): solve them at the same time, each on its own stream -- two thread-block
        # clusters of 16 SMs side by side.  Same kernels, same deterministic sums: the results are bit-identical either way.
        self.concurrent_uv = True
        self._pool = None
        self._streams = None
        self.log = []                       # (momentum_residual, continuity_residual, {'u','v','p': iters})

    # -- plumbing ---------------------------------------------------------------------------
    @property
    def _h(self):
        self._hd.use_current_stream()
        return self._hd.h

    def _params(self):
        P = _lib.SimpleParams()
        P.nx, P.ny = self.nx, self.ny
        P.dx, P.dy, P.dt, P.rho, P.mu = self.dx, self.dy, float(self.dt), self.rho, self.mu
        P.alpha_p, P.alpha_u, P.alpha_m = self.alpha_p, self.alpha_u, self.alpha_m
        for k, name in (('w', 'bc_w'), ('e', 'bc_e'), ('n', 'bc_n'), ('s', 'bc_s')):
            getattr(P, name)[0], getattr(P, name)[1] = float(self.bc[k][0]), float(self.bc[k][1])
        P.pin_p = 1 if self.variant == 'bicgsimp' else 0
        P.east_wall_bug = 0 if self.variant == 'bicgsimp' else 1
        P.pitch_u, P.pitch_v, P.pitch_p = self.u.pitch, self.v.pitch, self.p.pitch
        return P

    # -- CSV dumps (simple.py:51-75) ----------------------------------------------------------
    def dump_matrix(self, step, msg, directory='log'):
        os.makedirs(directory, exist_ok=True)
        for name, val in {'u': self.u, 'v': self.v, 'p': self.p, 'mdiv': self.mdiv, 'pcor': self.pcor}.items():
            np.savetxt(f'{directory}/{step:06}-{name}-{msg}.csv', val.to_numpy(), delimiter=',')

    def dump_coef(self, step, msg, directory='log'):
        os.makedirs(directory, exist_ok=True)
        for tag, coef, b in (('u', self.coef_u, self.b_u), ('v', self.coef_v, self.b_v), ('p', self.coef_p, self.b_p)):
            c = coef.to_numpy()
            for k, nm in enumerate(('ap', 'aw', 'ae', 'an', 'as')):
                np.savetxt(f'{directory}/{step:06}-{nm}{tag}-{msg}.csv', c[:, :, k], delimiter=',')
            np.savetxt(f'{directory}/{step:06}-b{tag} -{msg}.csv', b.to_numpy(), delimiter=',')

    # -- kernels -----------------------------------------------------------------------------------
    def compute_coef_u(self):
        """``bicgsimp.py:79-94`` -- fused with compute_coef_v in one launch; calling either assembles both."""
        self._assemble_uv()

    def compute_coef_v(self):
        """``bicgsimp.py:97-112``"""
        self._assemble_uv()

    def _assemble_uv(self):
        P = self._params()
        _lib.check(self._lib.kt_assemble_uv(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.p), _p(self.u0), _p(self.v0),
                                            _p(self.coef_u), _p(self.b_u), _p(self.coef_v), _p(self.b_v)))

    def compute_mdiv(self):
        """``bicgsimp.py:114-122``"""
        P = self._params()
        out = C.c_double()
        _lib.check(self._lib.kt_mdiv_max(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.mdiv), C.byref(out)))
        return out.value

    def compute_coef_p(self):
        """``bicgsimp.py:124-151``"""
        P = self._params()
        _lib.check(self._lib.kt_assemble_p(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.coef_u), _p(self.coef_v),
                                           _p(self.coef_p), _p(self.b_p), _p(self.mdiv)))

    def set_bc(self):
        """``bicgsimp.py:153-197``"""
        P = self._params()
        _lib.check(self._lib.kt_set_bc_cavity(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.coef_u), _p(self.b_u),
                                              _p(self.coef_v), _p(self.b_v)))

    def update_velocity(self):
        """``bicgsimp.py:212-225``"""
        P = self._params()
        out = C.c_double()
        _lib.check(self._lib.kt_relax_uv(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.u_mid), _p(self.v_mid), C.byref(out)))
        return out.value

    def correct_pressure(self):
        """``bicgsimp.py:232-236``"""
        P = self._params()
        _lib.check(self._lib.kt_correct_puv(self._h, C.byref(P), _p(self.p), None, None, _p(self.pcor), None, None))

    def correct_velocity(self):
        """``bicgsimp.py:238-244``"""
        P = self._params()
        _lib.check(self._lib.kt_correct_puv(self._h, C.byref(P), None, _p(self.u), _p(self.v), _p(self.pcor),
                                            _p(self.coef_u), _p(self.coef_v)))

    def correct_pressure_velocity(self):
        """correct_pressure + correct_velocity in one pass over pcor (what ``step()`` uses)."""
        P = self._params()
        _lib.check(self._lib.kt_correct_puv(self._h, C.byref(P), _p(self.p), _p(self.u), _p(self.v), _p(self.pcor),
                                            _p(self.coef_u), _p(self.coef_v)))

    def init(self):
        """``bicgsimp.py:246-256`` (``simple.py:257-264`` zeroes u, v, p only)."""
        fields = [self.u, self.v, self.p]
        if self.variant == 'bicgsimp':
            fields += [self.u0, self.v0, self.pcor]
        for f in fields:
            f.data.zero_()

    # -- drivers ------------------------------------------------------------------------------------
    def _make_solvers(self):
        self.u_momentum_solver = BICGSolver(self.coef_u, self.b_u, self.u_mid)
        self.v_momentum_solver = BICGSolver(self.coef_v, self.b_v, self.v_mid)
        self.p_correction_solver = CGSolver(self.coef_p, self.b_p, self.pcor)

    def bicg_solve_momentum_eqn(self, n_iter):
        """``bicgsimp.py:199-210``"""
        residual = 0.0
        if self.u_momentum_solver is None:
            self._make_solvers()
        for _ in range(n_iter):
            self.compute_coef_u()          # assembles u and v
            self.set_bc()
            self.u_momentum_solver.update_coef(self.coef_u, self.b_u, self.u_mid)
            self.v_momentum_solver.update_coef(self.coef_v, self.b_v, self.v_mid)
            if self.concurrent_uv:
                self._solve_uv_concurrently()
            else:
                self.u_momentum_solver.solve(eps=self.eps_u, quiet=True)
                self.v_momentum_solver.solve(eps=self.eps_v, quiet=True)
            residual = self.update_velocity()
        return residual

    def _solve_uv_concurrently(self):
        dev = self.p.device
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix='kt-v-momentum')
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))          # the assembly / BC kernels enqueued so far

        def run(solver, stream, eps):
            with torch.cuda.device(dev), torch.cuda.stream(stream):   # current device and stream are per thread
                stream.wait_event(ready)
                solver.solve(eps=eps, quiet=True)                      # returns when the solve has finished on `stream`
        fut = self._pool.submit(run, self.v_momentum_solver, self._streams[1], self.eps_v)
        try:
            run(self.u_momentum_solver, self._streams[0], self.eps_u)
        finally:
            fut.result()                                               # re-raises e.g. the reference's ZeroDivisionError

    solve_momentum_eqn = bicg_solve_momentum_eqn      # simple.py:243-250

    def bicg_solve_pcorrection_eqn(self, eps):
        """``bicgsimp.py:227-230``"""
        if self.p_correction_solver is None:
            self._make_solvers()
        self.compute_coef_p()
        self.p_correction_solver.update_coef(self.coef_p, self.b_p, self.pcor)
        self.p_correction_solver.solve(eps, quiet=True)

    def solve_pcorrection_eqn(self):                  # simple.py:252-255
        self.bicg_solve_pcorrection_eqn(self.eps_p)

    def step(self):
        """One SIMPLE outer iteration (``bicgsimp.py:313-317`` / ``simple.py:271-280``)."""
        momentum_residual = self.bicg_solve_momentum_eqn(1)
        if self.variant == 'simple':
            self.compute_mdiv()
        self.bicg_solve_pcorrection_eqn(self.eps_p)
        self.correct_pressure_velocity()
        continuity_residual = self.compute_mdiv()
        self.log.append((momentum_residual, continuity_residual,
                         {'u': self.u_momentum_solver.iterations, 'v': self.v_momentum_solver.iterations,
                          'p': self.p_correction_solver.iterations}))
        return momentum_residual, continuity_residual

    def solve(self, max_steps=None, init=True):
        """``bicgsimp.py:258-353`` / ``simple.py:266-285``.  Returns the number of outer iterations run."""
        if init:
            self.init()
        self._make_solvers()
        if max_steps is None:
            max_steps = 10000 if self.variant == 'bicgsimp' else 2000
        n = 0
        for substep in range(max_steps):
            momentum_residual, continuity_residual = self.step()
            n += 1
            if not self.quiet:
                print(f'>>> Solving step {substep:06} Current continuity residual: {continuity_residual:.3e} '
                      f'Current momentum residual: {momentum_residual:.3e}')
            if self.variant == 'bicgsimp' and momentum_residual < 1e-2 and continuity_residual < 1e-6:
                if not self.quiet:
                    print('>>> Solution converged.')
                break
        return n


    # -- checkpoint / restart (absent in the reference) -------------------------------------
    _STATE = ('u', 'v', 'p', 'u0', 'v0', 'pcor')

    def save_checkpoint(self, path):
        np.savez_compressed(path, nx=self.nx, ny=self.ny, lx=self.lx, ly=self.ly, variant=self.variant,
                            outer_iterations=len(self.log), **{k: getattr(self, k).to_numpy() for k in self._STATE})

    def load_checkpoint(self, path):
        z = np.load(path)
        if (int(z['nx']), int(z['ny'])) != (self.nx, self.ny):
            raise ValueError('checkpoint grid does not match this solver')
        for k in self._STATE:
            getattr(self, k).from_numpy(z[k])
        return int(z['outer_iterations'])


def lid_driven_cavity(n=50, lid=1.0, variant='bicgsimp', **kw):
    """The set-up of the reference script tail (``bicgsimp.py:356-366``: 50x50, lid speed 1)."""
    s = SIMPLESolver(1.0, 1.0, n, n, variant=variant, **kw)
    s.bc['n'][1] = lid
    return s
