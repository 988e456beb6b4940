"""``KarmanSolver`` -- the monolith ``karman_taichi.py`` (pressure-driven channel with a square-cell cylinder, the
README's Karman vortex street) on libkarman_b200.

The reference is a script of module-level fields and functions (``init``, ``fill_Au/fill_Av/fill_Ap``, ``bicgstab``,
``xu_back/xv_back/xp_back``, ``solve_momentum_bicgstab``, ``solve_pcorrection_bicgstab``, ``correct_conserv``,
``puv_correction``, ``time_forward`` and the main loop, ``karman_taichi.py:130-683``); this class keeps those names
as methods.  Its storage scheme (n x n pointer-SNode matrices, flattened 1-D vectors, one serial mega-kernel) is NOT
reproduced: every system lives in the class version's stencil arrays ``coef_*[i, j, 0..4]`` + a Jacobi plane, and is
solved by the same device-resident BiCGSTAB as everything else (``BICGSolver`` with ``Minv``, warm start and the
``omega == 0`` exit: ``karman_taichi.py:342-456``).  Importing this module runs nothing.

GUI / PNG output (``post_process_field``, ``ti.GUI``: ``:544-583,662-676``) is out of scope; ``write_matrix`` CSV dumps
are kept.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .bicgsolver import BICGSolver
from .cgsolver import _Handle, _p
from .field import Field


class KarmanSolver:
    def __init__(self, nx=320, ny=64, lx=0.5, ly=0.1, cyl=(64, 32, 144), dt=0.00001, momentum_iterations=1,
                 dtype=np.float64, device=None):
        # karman_taichi.py:7-23 (root file: 640x128; README / debug/karman_taichi.py: 320x64)
        self.nx, self.ny, self.lx, self.ly = nx, ny, lx, ly
        self.dx, self.dy = lx / nx, ly / ny
        self.rho, self.mu, self.dt = 1.0, 0.01, dt
        self.velo_rel, self.p_rel = 0.8, 0.2
        self.cyl = tuple(int(c) for c in cyl)
        self.momentum_iterations = momentum_iterations      # :636 (1); debug/karman_taichi.py:636 (30)
        self.real = np.dtype(dtype)
        mk = lambda *s: Field(s, self.real, device)
        self.p, self.pcor, self.udiv = mk(nx + 2, ny + 2), mk(nx + 2, ny + 2), mk(nx + 2, ny + 2)
        self.u, self.u0 = mk(nx + 3, ny + 2), mk(nx + 3, ny + 2)
        self.v, self.v0 = mk(nx + 2, ny + 3), mk(nx + 2, ny + 3)
        self.ct = torch.zeros((nx + 2, self.p.pitch), dtype=torch.int32, device=self.p.device)
        self.coef_u, self.b_u, self.M_u, self.Minv_u, self.x_u = mk(nx + 3, ny + 2, 5), mk(nx + 3, ny + 2), mk(nx + 3, ny + 2), mk(nx + 3, ny + 2), mk(nx + 3, ny + 2)
        self.coef_v, self.b_v, self.M_v, self.Minv_v, self.x_v = mk(nx + 2, ny + 3, 5), mk(nx + 2, ny + 3), mk(nx + 2, ny + 3), mk(nx + 2, ny + 3), mk(nx + 2, ny + 3)
        self.coef_p, self.b_p, self.M_p, self.Minv_p = mk(nx + 2, ny + 2, 5), mk(nx + 2, ny + 2), mk(nx + 2, ny + 2), mk(nx + 2, ny + 2)
        self.x_p = self.pcor                                 # xp_back is a copy of xp into pcor (:473-476): alias
        self._hd = _Handle(self.p)
        self._lib = self._hd.lib
        self.u_solver = self._jacobi_solver(self.coef_u, self.b_u, self.x_u, self.Minv_u, (nx + 1) * ny)
        self.v_solver = self._jacobi_solver(self.coef_v, self.b_v, self.x_v, self.Minv_v, nx * (ny + 1))
        self.p_solver = self._jacobi_solver(self.coef_p, self.b_p, self.x_p, self.Minv_p, nx * ny)
        self.last_iters = {'u': 0, 'v': 0, 'p': 0}
        self.last_status = {'u': 0, 'v': 0, 'p': 0}
        self.quiet = True

    @staticmethod
    def _jacobi_solver(coef, b, x, minv, n):
        s = BICGSolver(coef, b, x)
        s.Minv, s.warm_start, s.omega_zero_break = minv, True, True
        s.max_steps = 100 * n                                # :382
        return s

    @property
    def _h(self):
        self._hd.use_current_stream()
        return self._hd.h

    def _params(self):
        P = _lib.ChannelParams()
        P.nx, P.ny = self.nx, self.ny
        P.dx, P.dy, P.dt, P.rho, P.mu = self.dx, self.dy, float(self.dt), self.rho, self.mu
        P.velo_rel, P.p_rel = self.velo_rel, self.p_rel
        P.cyl_cx, P.cyl_cy, P.cyl_r2 = self.cyl
        P.pitch_u, P.pitch_v, P.pitch_p = self.u.pitch, self.v.pitch, self.p.pitch
        return P

    def _ct(self):
        return C.c_void_p(self.ct.data_ptr())

    def ct_numpy(self):
        return self.ct[:, :self.ny + 2].cpu().numpy()

    # -- reference functions ------------------------------------------------------------------------------
    def init(self):
        """``karman_taichi.py:130-154``"""
        for f in (self.pcor, self.x_u, self.x_v):
            f.data.zero_()
        P = self._params()
        _lib.check(self._lib.kt_channel_init(self._h, C.byref(P), _p(self.p), _p(self.u), _p(self.u0), _p(self.v), _p(self.v0), self._ct()))

    def fill_Au(self):
        """``karman_taichi.py:161-248`` -- assembled together with ``fill_Av`` in one launch."""
        self._assemble_uv()

    def fill_Av(self):
        """``karman_taichi.py:250-294``"""
        self._assemble_uv()

    def _assemble_uv(self):
        P = self._params()
        _lib.check(self._lib.kt_channel_assemble_uv(
            self._h, C.byref(P), self._ct(), _p(self.u), _p(self.v), _p(self.p), _p(self.u0), _p(self.v0),
            _p(self.coef_u), _p(self.b_u), _p(self.M_u), _p(self.Minv_u), _p(self.coef_v), _p(self.b_v), _p(self.M_v), _p(self.Minv_v)))

    def fill_Ap(self):
        """``karman_taichi.py:296-339``"""
        P = self._params()
        _lib.check(self._lib.kt_channel_assemble_p(self._h, C.byref(P), self._ct(), _p(self.u), _p(self.v), _p(self.coef_u),
                                                   _p(self.coef_v), _p(self.coef_p), _p(self.b_p), _p(self.M_p), _p(self.Minv_p)))

    def bicgstab(self, which, eps):
        """``karman_taichi.py:341-456`` for system 'u', 'v' or 'p'."""
        s = {'u': self.u_solver, 'v': self.v_solver, 'p': self.p_solver}[which]
        s.report_rho_zero = not self.quiet
        s.solve(eps=eps, quiet=True)
        self.last_iters[which], self.last_status[which] = s.iterations, s.status

    def xu_back(self):
        """``karman_taichi.py:458-463`` (+ ``xv_back`` ``:466-470`` in the same launch)"""
        P = self._params()
        _lib.check(self._lib.kt_channel_relax(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.x_u), _p(self.x_v)))

    xv_back = xu_back

    def xp_back(self):
        """``karman_taichi.py:473-476``: pcor aliases xp."""

    def solve_momentum_bicgstab(self, eps, max_iterations, solve_output=0, linalg_output=0):
        """``karman_taichi.py:480-494``"""
        for _ in range(max_iterations):
            self.fill_Au()
            self.bicgstab('u', eps)
            self.bicgstab('v', eps)
            self.xu_back()

    def solve_pcorrection_bicgstab(self, eps, linalg_output=0):
        """``karman_taichi.py:497-504``"""
        self.fill_Ap()
        self.bicgstab('p', eps)
        self.xp_back()

    def puv_correction(self):
        """``karman_taichi.py:506-542``; returns max |pcor|."""
        P = self._params()
        out = C.c_double()
        _lib.check(self._lib.kt_channel_correct(self._h, C.byref(P), self._ct(), _p(self.u), _p(self.v), _p(self.p), _p(self.pcor),
                                                _p(self.coef_u), _p(self.coef_v), C.byref(out)))
        return out.value

    def correct_conserv(self):
        """``karman_taichi.py:585-614``; returns (inlet sum, outlet sum before scaling, ratio)."""
        P = self._params()
        sums = (C.c_double * 3)()
        _lib.check(self._lib.kt_outlet_scale(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.udiv), sums))
        return tuple(sums)

    def time_forward(self):
        """``karman_taichi.py:616-622``"""
        self.u0.copy_from(self.u)
        self.v0.copy_from(self.v)

    def write_matrix(self, field, name):
        """``karman_taichi.py:157-159``"""
        np.savetxt(name + '.csv', field.to_numpy(), delimiter=',')

    # -- main loop (karman_taichi.py:625-660) ---------------------------------------------------------------
    def sub_iteration(self, eps=1e-8):
        self.solve_momentum_bicgstab(eps, self.momentum_iterations)
        self.correct_conserv()
        self.solve_pcorrection_bicgstab(eps)
        return self.puv_correction()

    def time_step(self, max_sub=500, eps=1e-8):
        pcor_max = 0.0
        n = 0
        for sub in range(max_sub):
            pcor_current = 10000.0
            if sub == 0:
                pcor_max = self.sub_iteration(eps)
            else:
                pcor_current = self.sub_iteration(eps)
            n += 1
            if not self.quiet:
                print(f'    >> The current pcor_max is {pcor_max: .3e} and the current p-correction compared with pcor_max is '
                      f'{(pcor_current / pcor_max): .3e}')
            if pcor_current < 1.0e-4 * pcor_max:
                break
        self.time_forward()
        return n

    # -- checkpoint / restart (the README run is 500 000 time steps, karman_taichi.py:628) ------
    _STATE = ('u', 'v', 'p', 'u0', 'v0', 'pcor', 'x_u', 'x_v')

    def save_checkpoint(self, path, time_step=0):
        np.savez_compressed(path, nx=self.nx, ny=self.ny, time_step=time_step, ct=self.ct_numpy(),
                            **{k: getattr(self, k).to_numpy() for k in self._STATE})

    def load_checkpoint(self, path):
        z = np.load(path)
        if (int(z['nx']), int(z['ny'])) != (self.nx, self.ny):
            raise ValueError('checkpoint grid does not match this solver')
        for k in self._STATE:
            getattr(self, k).from_numpy(z[k])
        self.ct[:, :self.ny + 2] = torch.from_numpy(z['ct']).to(self.ct.device)
        return int(z['time_step'])

    def run(self, time_steps, max_sub=500, eps=1e-8):
        self.init()
        return [self.time_step(max_sub, eps) for _ in range(time_steps)]
