"""T2 oracle: the monolith ``karman_taichi.py`` (channel + square-cell cylinder) restated in 5-point STENCIL form.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Vectorised NumPy; scales to the full 320x64 README case.  Proven
equal (bit for bit) to the literal matrix oracle ``oracle/karman_literal.py`` (T1) by ``tests/test_oracle_karman.py``.

The reference stores ``Au/Av/Ap`` as n x n sparse matrices over flattened unknowns (``karman_taichi.py:53-114``); every
entry it ever writes lies on the offsets ``0, +-1, +-stride``, so each system fits the class version's stencil arrays
``coef[i, j, 0..4] = (centre, i-1, i+1, j+1, j-1)`` (``cgsolver.py:59-60``) plus a separate Jacobi plane ``M``:

* u: unknowns ``u[i, j]``, i in [1, nx+1], j in [1, ny]   (``k = (i-1) ny + (j-1)``,   ``:165``)
* v: unknowns ``v[i, j]``, i in [1, nx],   j in [1, ny+1] (``k = (i-1)(ny+1) + (j-1)``, ``:254``)
* p: unknowns ``pcor[i, j]``, i in [1, nx], j in [1, ny]  (``k = (i-1) ny + (j-1)``,   ``:300``)

``k+-1`` is the j+-1 neighbour (slots 3 / 4), ``k+-stride`` the i+-1 neighbour (slots 2 / 1).  Reference quirks that
change the numbers are reproduced: first-match ``if/elif`` order, the second u-loop that rewrites rows
next to the obstacle WITHOUT clearing the stale ``j-1`` entry and WITHOUT refreshing ``Mu`` (``:221-239``), the dead third
branch (``:241``), outlet pressure rows whose diagonal is set to 1 while their off-diagonals stay (``:327-339``), the row
of cell (nx, 1) left un-pinned.  ``ct``: 1 solid, -1 fluid (``:143-150``).
"""
from __future__ import annotations

import numpy as np

from . import krylov


class KarmanOracle:
    def __init__(self, nx=320, ny=64, lx=0.5, ly=0.1, cyl=(64, 32, 144), dt=0.00001, momentum_iterations=1):
        self.nx, self.ny, self.lx, self.ly = nx, ny, lx, ly
        self.dx, self.dy = lx / nx, ly / ny
        self.rho, self.mu, self.dt = 1.0, 0.01, dt
        self.velo_rel, self.p_rel = 0.8, 0.2
        self.cyl = cyl
        self.momentum_iterations = momentum_iterations
        z = lambda *s: np.zeros(s)
        self.p, self.pcor, self.udiv = z(nx + 2, ny + 2), z(nx + 2, ny + 2), z(nx + 2, ny + 2)
        self.u, self.u0 = z(nx + 3, ny + 2), z(nx + 3, ny + 2)
        self.v, self.v0 = z(nx + 2, ny + 3), z(nx + 2, ny + 3)
        self.ct = np.zeros((nx + 2, ny + 2), dtype=np.int32)
        self.coef_u, self.b_u, self.M_u, self.x_u = z(nx + 3, ny + 2, 5), z(nx + 3, ny + 2), z(nx + 3, ny + 2), z(nx + 3, ny + 2)
        self.coef_v, self.b_v, self.M_v, self.x_v = z(nx + 2, ny + 3, 5), z(nx + 2, ny + 3), z(nx + 2, ny + 3), z(nx + 2, ny + 3)
        self.coef_p, self.b_p, self.M_p, self.x_p = z(nx + 2, ny + 2, 5), z(nx + 2, ny + 2), z(nx + 2, ny + 2), z(nx + 2, ny + 2)
        self.last_iters = {'u': 0, 'v': 0, 'p': 0}
        self.last_status = {'u': 0, 'v': 0, 'p': 0}

    # karman_taichi.py:130-154
    def init(self):
        nx, ny = self.nx, self.ny
        i = np.arange(nx + 2, dtype=np.float64)
        self.p[:] = (100 - 4000.0 * i / nx)[:, None]
        self.u[:] = 0.0; self.u0[:] = 0.0; self.v[:] = 0.0; self.v0[:] = 0.0
        self.ct[:] = 1
        self.ct[1:nx + 1, 1:ny + 1] = -1
        cx, cy, r2 = self.cyl
        ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing='ij')
        self.ct[:nx, :ny][((ii - cx) ** 2 + (jj - cy) ** 2) < r2] = 1

    # karman_taichi.py:161-248
    def fill_Au(self):
        nx, ny, dx, dy, dt, rho, mu = self.nx, self.ny, self.dx, self.dy, self.dt, self.rho, self.mu
        ct, u, v, p, u0 = self.ct, self.u, self.v, self.p, self.u0
        ic, im, ip = slice(1, nx + 2), slice(0, nx + 1), slice(2, nx + 3)
        jc, jm, jp = slice(1, ny + 1), slice(0, ny), slice(2, ny + 2)
        ct_l, ct_r = ct[im, jc], ct[ic, jc]
        ct_ru, ct_rd, ct_lu = ct[ic, jm], ct[ic, jp], ct[im, jm]
        aN = -mu * dx / dy - rho * 0.5 * (v[im, jp] + v[ic, jp]) * dx
        aS = -mu * dx / dy - (-rho * 0.5 * (v[im, jc] + v[ic, jc]) * dx)
        aW = -mu * dy / dx - rho * 0.5 * (u[ic, jc] + u[im, jc]) * dy
        aE = -mu * dy / dx - (-rho * 0.5 * (u[ic, jc] + u[ip, jc]) * dy)
        tt = rho * dx * dy / dt
        bb = (p[im, jc] - p[ic, jc]) * dy + rho * dx * dy / dt * u0[ic, jc]
        C_e = -aS - aN - aW - aE + tt
        C_c = -aN - aW - aE + tt + 2 * mu * dx / dy
        C_d = -aS - aW - aE + tt + 2 * mu * dx / dy
        m_a = ct_l == 1
        m_b = ~m_a & (ct_r == 1)
        rest = ~m_a & ~m_b
        m_c = rest & ((ct_r + ct_ru) == 0)
        m_d = rest & ~m_c & ((ct_r + ct_rd) == 0)
        m_e = rest & ~m_c & ~m_d
        shape = aN.shape
        C, W, E, N, S, B = (np.zeros(shape) for _ in range(6))
        C[m_a] = 1.0; E[m_a] = -1.0
        C[m_b] = 1.0; W[m_b] = -1.0
        for m, cc, has_n, has_s in ((m_c, C_c, True, False), (m_d, C_d, False, True), (m_e, C_e, True, True)):
            C[m] = cc[m]; W[m] = aW[m]; E[m] = aE[m]; B[m] = bb[m]
            if has_n:
                N[m] = aN[m]
            if has_s:
                S[m] = aS[m]
        M = C.copy()                                       # Mu[k,k] = Au[k,k] after loop 1 only (:218)
        # second loop (:221-248): i in [2, nx], j in [2, ny-1]
        region = np.zeros(shape, bool)
        region[1:nx, 1:ny - 1] = True                      # local index = i - 1, j - 1
        s2 = ct_l + ct_r
        m_A = region & ((s2 == 0) | (s2 == 2))
        m_B = region & ~m_A & (((ct_r + ct_ru) == 0) | ((ct_l + ct_lu) == 0))
        W[m_A] = 0.0; E[m_A] = 0.0; N[m_A] = 0.0; S[m_A] = 0.0; C[m_A] = 1.0; B[m_A] = 0.0
        N[m_B] = aN[m_B]; W[m_B] = aW[m_B]; E[m_B] = aE[m_B]; C[m_B] = C_c[m_B]; B[m_B] = bb[m_B]   # S keeps loop 1's value
        c = self.coef_u
        c[ic, jc, 0], c[ic, jc, 1], c[ic, jc, 2], c[ic, jc, 3], c[ic, jc, 4] = C, W, E, N, S
        self.b_u[ic, jc] = B
        self.M_u[ic, jc] = M

    # karman_taichi.py:250-294
    def fill_Av(self):
        nx, ny, dx, dy, dt, rho, mu = self.nx, self.ny, self.dx, self.dy, self.dt, self.rho, self.mu
        ct, u, v, p, v0 = self.ct, self.u, self.v, self.p, self.v0
        ic, im, ip = slice(1, nx + 1), slice(0, nx), slice(2, nx + 2)
        jc, jm, jp = slice(1, ny + 2), slice(0, ny + 1), slice(2, ny + 3)
        ct_c, ct_u = ct[ic, jc], ct[ic, jm]
        ct_w, ct_e = ct[im, jc], ct[ip, jc]
        aS = -mu * dx / dy - (-rho * 0.5 * (v[ic, jm] + v[ic, jc]) * dx)
        aN = -mu * dx / dy - (rho * 0.5 * (v[ic, jp] + v[ic, jc]) * dx)
        aE = -mu * dy / dx - (-rho * 0.5 * (u[ip, jm] + u[ip, jc]) * dy)
        aW = -mu * dy / dx - (rho * 0.5 * (u[ic, jc] + u[ic, jm]) * dy)
        tt = rho * dx * dy / dt
        bb = (p[ic, jc] - p[ic, jm]) * dx + rho * dx * dy / dt * v0[ic, jc]
        C_b = -aS - aN - aE + tt + 2 * mu * dy / dx
        C_c = -aS - aN - aW + tt + 2 * mu * dy / dx
        C_d = -aS - aN - aW - aE + tt
        s = ct_c + ct_u
        m_a = (s == 0) | (s == 2)
        m_b = ~m_a & ((ct_c + ct_w) == 0)
        m_c = ~m_a & ~m_b & ((ct_c + ct_e) == 0)
        m_d = ~m_a & ~m_b & ~m_c
        shape = aS.shape
        C, W, E, N, S, B = (np.zeros(shape) for _ in range(6))
        C[m_a] = 1.0; B[m_a] = v[ic, jc][m_a]
        for m, cc, has_w, has_e in ((m_b, C_b, False, True), (m_c, C_c, True, False), (m_d, C_d, True, True)):
            C[m] = cc[m]; S[m] = aS[m]; N[m] = aN[m]; B[m] = bb[m]
            if has_w:
                W[m] = aW[m]
            if has_e:
                E[m] = aE[m]
        c = self.coef_v
        c[ic, jc, 0], c[ic, jc, 1], c[ic, jc, 2], c[ic, jc, 3], c[ic, jc, 4] = C, W, E, N, S
        self.b_v[ic, jc] = B
        self.M_v[ic, jc] = C                               # Mv = diag(Av) (:292-294)

    # karman_taichi.py:296-339
    def fill_Ap(self):
        nx, ny, dx, dy, rho = self.nx, self.ny, self.dx, self.dy, self.rho
        ct, u, v = self.ct, self.u, self.v
        ic, im, ip = slice(1, nx + 1), slice(0, nx), slice(2, nx + 2)
        jc, jm, jp = slice(1, ny + 1), slice(0, ny), slice(2, ny + 2)
        Cu, Cv = self.coef_u[:, :, 0], self.coef_v[:, :, 0]
        B = rho * (u[ic, jc] - u[ip, jc]) * dy + rho * (v[ic, jp] - v[ic, jc]) * dx
        with np.errstate(divide='ignore', invalid='ignore'):
            S = -rho * dx * dx / Cv[ic, jc]
            N = -rho * dx * dx / Cv[ic, jp]
            W = -rho * dy * dy / Cu[ic, jc]
            E = -rho * dy * dy / Cu[ip, jc]
        ct_c = ct[ic, jc]
        S[(ct_c + ct[ic, jm]) == 0] = 0.0
        N[(ct_c + ct[ic, jp]) == 0] = 0.0
        W[(ct_c + ct[im, jc]) == 0] = 0.0
        E[(ct_c + ct[ip, jc]) == 0] = 0.0
        # entries the reference writes outside the matrix are dropped and read back as 0 (k-1 < 0, k-ny < 0, k+ny >= n)
        S[0, 0] = 0.0
        W[0, :] = 0.0
        E[nx - 1, :] = 0.0
        C = -S - N - W - E
        M = C.copy()
        # second loop (:327-339): rows 0..ny-1 -> identity; rows n-1..n-ny+1 -> diagonal 1, b 0 (off-diagonals stay)
        C[0, :] = 1.0; M[0, :] = 1.0; B[0, :] = 0.0
        S[0, :] = 0.0; N[0, :] = 0.0; W[0, :] = 0.0; E[0, :] = 0.0
        C[nx - 1, 1:] = 1.0; M[nx - 1, 1:] = 1.0; B[nx - 1, 1:] = 0.0
        c = self.coef_p
        c[ic, jc, 0], c[ic, jc, 1], c[ic, jc, 2], c[ic, jc, 3], c[ic, jc, 4] = C, W, E, N, S
        self.b_p[ic, jc] = B
        self.M_p[ic, jc] = M

    def _solve(self, which, coef, M, b, x, eps, n):
        xs, it, st = krylov.jacobi_bicgstab(coef, M, b, x, eps=eps, n=n, max_iter=getattr(self, 'max_iter', None))
        x[:] = xs
        self.last_iters[which], self.last_status[which] = it, st

    # karman_taichi.py:458-504
    def solve_momentum_bicgstab(self, eps, max_iterations):
        nx, ny = self.nx, self.ny
        for _ in range(max_iterations):
            self.fill_Au()
            self.fill_Av()
            self._solve('u', self.coef_u, self.M_u, self.b_u, self.x_u, eps, (nx + 1) * ny)
            self._solve('v', self.coef_v, self.M_v, self.b_v, self.x_v, eps, nx * (ny + 1))
            U = self.u[1:nx + 2, 1:ny + 1]
            U[:] = U + self.velo_rel * (self.x_u[1:nx + 2, 1:ny + 1] - U)          # xu_back
            V = self.v[1:nx + 1, 1:ny + 2]
            V[:] = V + self.velo_rel * (self.x_v[1:nx + 1, 1:ny + 2] - V)          # xv_back

    def solve_pcorrection_bicgstab(self, eps):
        nx, ny = self.nx, self.ny
        self.fill_Ap()
        self._solve('p', self.coef_p, self.M_p, self.b_p, self.x_p, eps, nx * ny)
        self.pcor[1:nx + 1, 1:ny + 1] = self.x_p[1:nx + 1, 1:ny + 1]              # xp_back

    # karman_taichi.py:506-542
    def puv_correction(self):
        nx, ny, dx, dy = self.nx, self.ny, self.dx, self.dy
        ct, pcor = self.ct, self.pcor
        ic, im = slice(1, nx + 2), slice(0, nx + 1)
        jc = slice(1, ny + 1)
        s = ct[im, jc] + ct[ic, jc]
        m = ~((s == 0) | (s == 2))
        with np.errstate(all='ignore'):
            ucor = (pcor[im, jc] - pcor[ic, jc]) * dy / self.coef_u[ic, jc, 0]
        U = self.u[ic, jc]
        U[m] = U[m] + ucor[m]
        ic, jc, jm = slice(1, nx + 1), slice(1, ny + 2), slice(0, ny + 1)
        s = ct[ic, jc] + ct[ic, jm]
        m = ~((s == 0) | (s == 2))
        with np.errstate(all='ignore'):
            vcor = (pcor[ic, jc] - pcor[ic, jm]) * dx / self.coef_v[ic, jc, 0]
        V = self.v[ic, jc]
        V[m] = V[m] + vcor[m]
        fluid = ct != 1
        self.p[fluid] = self.p[fluid] + self.p_rel * pcor[fluid]
        return float(np.max(np.abs(pcor)))

    # karman_taichi.py:585-614
    def correct_conserv(self):
        nx, ny, dx, dy = self.nx, self.ny, self.dx, self.dy
        u, v = self.u, self.v
        mdot_inlet = float(krylov._sum(u[1, 1:ny + 1], 'column'))
        mdot_outlet = float(krylov._sum(u[nx + 1, 1:ny + 1], 'column'))
        with np.errstate(all='ignore'):
            coef = np.float64(mdot_inlet) / np.float64(mdot_outlet)
        u[nx + 1, 1:ny + 1] = coef * u[nx + 1, 1:ny + 1]
        self.udiv[1:nx + 1, 1:ny + 1] = (u[1:nx + 1, 1:ny + 1] - u[2:nx + 2, 1:ny + 1]) * dy + \
            (v[1:nx + 1, 2:ny + 2] - v[1:nx + 1, 1:ny + 1]) * dx

    def time_forward(self):
        self.u0[:] = self.u
        self.v0[:] = self.v

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
            if pcor_current < 1.0e-4 * pcor_max:
                break
        self.time_forward()
        return n
