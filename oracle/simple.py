"""NumPy restatement of the reference's class-based SIMPLE solver (lid-driven cavity).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows ``bicgsimp.py:11-353`` (the
semantic default: the version the reference validates against ``data/ghia1982.dat``)
with ``simple.py:11-285`` as ``variant='simple'`` (differences: SURVEY.md A.3).

Staggered layout (``bicgsimp.py:27-47``): ``u (nx+3, ny+2)`` on west faces, unknowns
i in [2,nx], j in [1,ny]; ``v (nx+2, ny+3)`` on south faces, unknowns i in [1,nx],
j in [2,ny]; ``p, pcor, mdiv (nx+2, ny+2)``; ``coef_* (..., 5)`` = (ap, aw, ae, an, as).

Every expression keeps the reference's multiplication/addition order (Python and
Taichi both evaluate the written expression left to right) and NumPy does not contract
to FMA, so a ``-fmad=false`` device kernel reproduces these arrays bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

from . import krylov


class SimpleOracle:
    def __init__(self, lx, ly, nx, ny, variant='bicgsimp', dtype=np.float64):
        assert variant in ('bicgsimp', 'simple')
        self.variant = variant
        self.lx, self.ly, self.nx, self.ny = lx, ly, nx, ny
        self.dx = self.lx / self.nx
        self.dy = self.ly / self.ny
        self.rho = 1.00
        self.mu = 0.01
        if variant == 'bicgsimp':           # bicgsimp.py:20-25
            self.dt = 1e12
            self.alpha_p, self.alpha_u, self.alpha_m = 0.1, 0.8, 0.05
            self.eps_u = self.eps_v = 1e-4  # bicgsimp.py:206,208
            self.eps_p = 1e-8               # bicgsimp.py:314
        else:                               # simple.py:20-25
            self.dt = 10000
            self.alpha_p, self.alpha_u, self.alpha_m = 0.001, 0.001, 0.001
            self.eps_u = self.eps_v = 1e-1  # simple.py:248-249
            self.eps_p = 1e-2               # simple.py:255
        self.dtype = dtype
        z = lambda *s: np.zeros(s, dtype=dtype)
        self.u, self.u_mid, self.u0 = z(nx + 3, ny + 2), z(nx + 3, ny + 2), z(nx + 3, ny + 2)
        self.v, self.v_mid, self.v0 = z(nx + 2, ny + 3), z(nx + 2, ny + 3), z(nx + 2, ny + 3)
        self.p, self.pcor, self.mdiv = z(nx + 2, ny + 2), z(nx + 2, ny + 2), z(nx + 2, ny + 2)
        self.bc = {'w': [0.0, 0.0], 'e': [0.0, 0.0], 'n': [0.0, 0.0], 's': [0.0, 0.0]}
        self.coef_u, self.b_u = z(nx + 3, ny + 2, 5), z(nx + 3, ny + 2)
        self.coef_v, self.b_v = z(nx + 2, ny + 3, 5), z(nx + 2, ny + 3)
        self.coef_p, self.b_p = z(nx + 2, ny + 2, 5), z(nx + 2, ny + 2)
        # bookkeeping (not in the reference): Krylov iterations / status of the last solves
        self.last_iters = {'u': 0, 'v': 0, 'p': 0}
        self.last_status = {'u': 0, 'v': 0, 'p': 0}
        self.log = []

    # -- assembly ---------------------------------------------------------------------------
    def compute_coef_u(self):
        """``bicgsimp.py:79-94`` (= ``simple.py:77-92``), i in [2,nx], j in [1,ny]."""
        nx, ny, dx, dy, dt, rho, mu = self.nx, self.ny, self.dx, self.dy, self.dt, self.rho, self.mu
        u, v, p, u0, c = self.u, self.v, self.p, self.u0, self.coef_u
        I, J = slice(2, nx + 1), slice(1, ny + 1)
        uc, uw, ue = u[I, J], u[1:nx, J], u[3:nx + 2, J]
        v_wn, v_n = v[1:nx, 2:ny + 2], v[I, 2:ny + 2]      # v[i-1,j+1], v[i,j+1]
        v_ws, v_s = v[1:nx, J], v[I, J]                    # v[i-1,j],   v[i,j]
        c[I, J, 1] = -(mu * dy / dx + 0.5 * rho * 0.5 * (uc + uw) * dy)
        c[I, J, 2] = -(mu * dy / dx - 0.5 * rho * 0.5 * (uc + ue) * dy)
        c[I, J, 3] = -(mu * dx / dy - 0.5 * rho * 0.5 * (v_wn + v_n) * dx)
        c[I, J, 4] = -(mu * dx / dy + 0.5 * rho * 0.5 * (v_ws + v_s) * dx)
        c[I, J, 0] = -(c[I, J, 1] + c[I, J, 2] + c[I, J, 3] + c[I, J, 4]) + \
            rho * 0.5 * (uc + ue) * dy - \
            rho * 0.5 * (uc + uw) * dy + \
            rho * 0.5 * (v_wn + v_n) * dx - \
            rho * 0.5 * (v_ws + v_s) * dx + \
            rho * dx * dy / dt
        self.b_u[I, J] = (p[1:nx, J] - p[I, J]) * dy + rho * dx * dy / dt * u0[I, J]

    def compute_coef_v(self):
        """``bicgsimp.py:97-112`` (= ``simple.py:95-110``), i in [1,nx], j in [2,ny]."""
        nx, ny, dx, dy, dt, rho, mu = self.nx, self.ny, self.dx, self.dy, self.dt, self.rho, self.mu
        u, v, p, v0, c = self.u, self.v, self.p, self.v0, self.coef_v
        I, J = slice(1, nx + 1), slice(2, ny + 1)
        vc, vn, vs = v[I, J], v[I, 3:ny + 2], v[I, 1:ny]
        u_c, u_s = u[I, J], u[I, 1:ny]                      # u[i,j],   u[i,j-1]
        u_es, u_e = u[2:nx + 2, 1:ny], u[2:nx + 2, J]       # u[i+1,j-1], u[i+1,j]
        c[I, J, 1] = -(mu * dy / dx + 0.5 * rho * 0.5 * (u_c + u_s) * dy)
        c[I, J, 2] = -(mu * dy / dx - 0.5 * rho * 0.5 * (u_es + u_e) * dy)
        c[I, J, 3] = -(mu * dx / dy - 0.5 * rho * 0.5 * (vn + vc) * dx)
        c[I, J, 4] = -(mu * dx / dy + 0.5 * rho * 0.5 * (vs + vc) * dx)
        c[I, J, 0] = -(c[I, J, 1] + c[I, J, 2] + c[I, J, 3] + c[I, J, 4]) + \
            rho * 0.5 * (u_es + u_e) * dy - \
            rho * 0.5 * (u_c + u_s) * dy + \
            rho * 0.5 * (vn + vc) * dx - \
            rho * 0.5 * (vs + vc) * dx + \
            rho * dx * dy / dt
        self.b_v[I, J] = (p[I, 1:ny] - p[I, J]) * dx + rho * dx * dy / dt * v0[I, J]

    def _mdiv_expr(self):
        nx, ny, dx, dy, rho = self.nx, self.ny, self.dx, self.dy, self.rho
        u, v = self.u, self.v
        I, J = slice(1, nx + 1), slice(1, ny + 1)
        return rho * (u[I, J] - u[2:nx + 2, J]) * dy + rho * (v[I, J] - v[I, 2:ny + 2]) * dx

    def compute_mdiv(self):
        """``bicgsimp.py:114-122``; returns max |mdiv| (``simple.py:137-141`` returns None)."""
        nx, ny = self.nx, self.ny
        self.mdiv[1:nx + 1, 1:ny + 1] = self._mdiv_expr()
        return float(np.max(np.abs(self.mdiv[1:nx + 1, 1:ny + 1]), initial=0.0))

    def compute_coef_p(self):
        """``bicgsimp.py:124-151`` (pin active) / ``simple.py:143-170`` (pin commented out)."""
        nx, ny, dx, dy, rho = self.nx, self.ny, self.dx, self.dy, self.rho
        cu, cv, c = self.coef_u, self.coef_v, self.coef_p
        I, J = slice(1, nx + 1), slice(1, ny + 1)
        self.mdiv[I, J] = self._mdiv_expr()
        self.b_p[I, J] = self.mdiv[I, J]
        with np.errstate(divide='ignore', invalid='ignore'):
            c[I, J, 1] = -rho * dy * dy / cu[I, J, 0]
            c[I, J, 2] = -rho * dy * dy / cu[2:nx + 2, J, 0]
            c[I, J, 3] = -rho * dx * dx / cv[I, 2:ny + 2, 0]
            c[I, J, 4] = -rho * dx * dx / cv[I, J, 0]
        c[1, J, 1] = 0.0
        c[nx, J, 2] = 0.0
        c[I, 1, 4] = 0.0
        c[I, ny, 3] = 0.0
        c[I, J, 0] = -(c[I, J, 1] + c[I, J, 2] + c[I, J, 3] + c[I, J, 4])
        if self.variant == 'bicgsimp':      # bicgsimp.py:146-151
            c[1, 1, 1:5] = 0.0
            c[1, 1, 0] = 1.0
            self.b_p[1, 1] = 0.0

    def set_bc(self):
        """``bicgsimp.py:153-197``: four sequential 1-D loops, statement order kept."""
        nx, ny, bc, mu, dx, dy = self.nx, self.ny, self.bc, self.mu, self.dx, self.dy
        cu, cv, bu, bv, u, v = self.coef_u, self.coef_v, self.b_u, self.b_v, self.u, self.v
        J = slice(1, ny + 1)
        bu[2, J] += -cu[2, J, 1] * bc['w'][0]
        cu[2, J, 1] = 0.0
        u[1, J] = bc['w'][0]
        bu[nx, J] += -cu[nx, J, 2] * bc['e'][0]
        cu[nx, J, 2] = 0.0
        u[nx + 1, J] = bc['e'][0]

        I = slice(1, nx + 1)
        bv[I, 2] += -cv[I, 2, 4] * bc['s'][0]
        cv[I, 2, 4] = 0.0
        v[I, 1] = bc['s'][0]
        bv[I, ny] += -cv[I, ny, 3] * bc['n'][0]
        cv[I, ny, 3] = 0.0
        v[I, ny + 1] = bc['n'][0]

        I = slice(2, nx + 1)
        bu[I, 1] += 2 * mu * bc['s'][1] * dx / dy
        cu[I, 1, 0] += (cu[I, 1, 4] + 2 * mu * dx / dy)
        cu[I, 1, 4] = 0.0
        bu[I, ny] += 2 * mu * bc['n'][1] * dx / dy
        cu[I, ny, 0] += (cu[I, ny, 3] + 2 * mu * dx / dy)
        cu[I, ny, 3] = 0.0

        J = slice(2, ny + 1)
        bv[1, J] += 2 * mu * bc['w'][1] * dy / dx
        cv[1, J, 0] += (cv[1, J, 1] + 2 * mu * dy / dx)
        cv[1, J, 1] = 0.0
        bv[nx, J] += 2 * mu * bc['e'][1] * dy / dx
        k_east = 2 if self.variant == 'bicgsimp' else 1   # simple.py:215 adds aw (slot 1): bug kept
        cv[nx, J, 0] += (cv[nx, J, k_east] + 2 * mu * dy / dx)
        cv[nx, J, 2] = 0.0

    # -- relaxation / corrections -----------------------------------------------------------------
    def update_velocity(self):
        """``bicgsimp.py:212-225`` (max |diff| taken before the blend); ``simple.py:232-238``."""
        nx, ny, am = self.nx, self.ny, self.alpha_m
        I, J = slice(2, nx + 1), slice(1, ny + 1)
        max_udiff = float(np.max(np.abs(self.u_mid[I, J] - self.u[I, J]), initial=0.0))
        self.u[I, J] = am * self.u_mid[I, J] + (1 - am) * self.u[I, J]
        I, J = slice(1, nx + 1), slice(2, ny + 1)
        max_vdiff = float(np.max(np.abs(self.v_mid[I, J] - self.v[I, J]), initial=0.0))
        self.v[I, J] = am * self.v_mid[I, J] + (1 - am) * self.v[I, J]
        return math.sqrt(max_udiff ** 2 + max_vdiff ** 2)

    def correct_pressure(self):
        """``bicgsimp.py:232-236``."""
        nx, ny = self.nx, self.ny
        I, J = slice(1, nx + 1), slice(1, ny + 1)
        self.p[I, J] += self.alpha_p * self.pcor[I, J]

    def correct_velocity(self):
        """``bicgsimp.py:238-244``."""
        nx, ny, dx, dy, au = self.nx, self.ny, self.dx, self.dy, self.alpha_u
        pc = self.pcor
        I, J = slice(2, nx + 1), slice(1, ny + 1)
        self.u[I, J] += au * (pc[1:nx, J] - pc[I, J]) * dy / self.coef_u[I, J, 0]
        I, J = slice(1, nx + 1), slice(2, ny + 1)
        self.v[I, J] += au * (pc[I, 1:ny] - pc[I, J]) * dx / self.coef_v[I, J, 0]

    def init(self):
        """``bicgsimp.py:246-256`` (``simple.py:257-264`` zeroes u, v, p only)."""
        self.u[:] = 0.0
        self.v[:] = 0.0
        self.p[:] = 0.0
        if self.variant == 'bicgsimp':
            self.u0[:] = 0.0
            self.v0[:] = 0.0
            self.pcor[:] = 0.0

    # -- drivers ------------------------------------------------------------------------------
    def _krylov(self, which, kind, coef, b, eps):
        fn = krylov.bicgstab_solve if kind == 'bicg' else krylov.cg_solve
        x, it, st = fn(coef, b, eps=eps)
        self.last_iters[which], self.last_status[which] = it, st
        return x

    def solve_momentum_eqn(self, n_iter=1):
        """``bicgsimp.py:199-210`` / ``simple.py:243-250``."""
        residual = 0.0
        for _ in range(n_iter):
            self.compute_coef_u()
            self.compute_coef_v()
            self.set_bc()
            self.u_mid[:] = self._krylov('u', 'bicg', self.coef_u, self.b_u, self.eps_u)
            self.v_mid[:] = self._krylov('v', 'bicg', self.coef_v, self.b_v, self.eps_v)
            residual = self.update_velocity()
        return residual

    def solve_pcorrection_eqn(self, eps=None):
        """``bicgsimp.py:227-230`` / ``simple.py:252-255``."""
        self.compute_coef_p()
        self.pcor[:] = self._krylov('p', 'cg', self.coef_p, self.b_p, self.eps_p if eps is None else eps)

    def step(self):
        """One SIMPLE outer iteration; returns (momentum_residual, continuity_residual)."""
        if self.variant == 'bicgsimp':       # bicgsimp.py:313-317
            mres = self.solve_momentum_eqn(1)
            self.solve_pcorrection_eqn()
            self.correct_pressure()
            self.correct_velocity()
            cres = self.compute_mdiv()
        else:                                # simple.py:271-280
            mres = self.solve_momentum_eqn(1)
            self.compute_mdiv()
            self.solve_pcorrection_eqn()
            self.correct_pressure()
            self.correct_velocity()
            cres = self.compute_mdiv()
        self.log.append((mres, cres, dict(self.last_iters)))
        return mres, cres

    def solve(self, max_steps=None, init=True):
        """``bicgsimp.py:258-353`` (stop: momentum < 1e-2 and continuity < 1e-6, ``:326``) or
        ``simple.py:266-285`` (fixed 2000 steps).  Returns the number of outer iterations run."""
        if init:
            self.init()
        if max_steps is None:
            max_steps = 10000 if self.variant == 'bicgsimp' else 2000
        n = 0
        for _ in range(max_steps):
            mres, cres = self.step()
            n += 1
            if self.variant == 'bicgsimp' and mres < 1e-2 and cres < 1e-6:
                break
        return n
