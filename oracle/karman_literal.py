"""T1 oracle: LITERAL restatement of the monolith ``karman_taichi.py`` (channel + square-cell cylinder).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Trust anchor for the stencil-form oracle ``oracle/karman.py`` (T2):
the matrices ``Au/Av/Ap`` and ``Mu/Mv/Mp`` are filled with the reference's own index arithmetic
(``karman_taichi.py:162-339``) into dict-of-keys "pointer SNode" stand-ins, one Python loop iteration per reference loop
iteration, first-match ``if/elif`` chains kept verbatim.  Semantics assumed for Taichi's sparse fields:
entries persist between fills (nothing ever clears them), reads of never-written or out-of-range entries give 0,
out-of-range writes are dropped.  Small grids only (pure-Python loops).

The BiCGSTAB kernel (``:342-456``) visits the 7 offsets ``[i-ny-1, i-ny, i-1, i, i+1, i+ny, i+ny+1]`` of each row; here the
dict is converted to CSR restricted to exactly those offsets (``to_csr7``) so the Krylov loop runs at NumPy speed.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

from . import krylov


class SparseMat:
    """n x n pointer-SNode matrix stand-in."""

    def __init__(self, n):
        self.n = n
        self.d = {}

    def __setitem__(self, ij, val):
        i, j = ij
        if 0 <= i < self.n and 0 <= j < self.n:
            self.d[(i, j)] = float(val)

    def __getitem__(self, ij):
        return self.d.get(ij, 0.0)

    def to_csr7(self, ny):
        """CSR over the 7 offsets the reference's product visits; also returns the entries it would NOT visit."""
        offs = {-ny - 1, -ny, -1, 0, 1, ny, ny + 1}
        rows, cols, vals, missed = [], [], [], []
        for (i, j), v in self.d.items():
            if (j - i) in offs:
                rows.append(i); cols.append(j); vals.append(v)
            elif v != 0.0:
                missed.append((i, j, v))
        return sp.csr_matrix((vals, (rows, cols)), shape=(self.n, self.n)), missed

    def diag(self):
        return np.array([self.d.get((k, k), 0.0) for k in range(self.n)])


class KarmanLiteral:
    def __init__(self, nx=320, ny=64, lx=0.5, ly=0.1, cyl=(64, 32, 144), dt=0.00001, momentum_iterations=1):
        self.nx, self.ny, self.lx, self.ly = nx, ny, lx, ly
        self.dx, self.dy = lx / nx, ly / ny
        self.rho, self.mu, self.dt = 1.0, 0.01, dt
        self.velo_rel, self.p_rel = 0.8, 0.2
        self.cyl = cyl
        self.momentum_iterations = momentum_iterations     # karman_taichi.py:636 (1) / debug/karman_taichi.py:636 (30)
        z = lambda *s: np.zeros(s)
        self.p, self.pcor, self.udiv = z(nx + 2, ny + 2), z(nx + 2, ny + 2), z(nx + 2, ny + 2)
        self.u, self.u0, self.ucor = z(nx + 3, ny + 2), z(nx + 3, ny + 2), z(nx + 3, ny + 2)
        self.v, self.v0, self.vcor = z(nx + 2, ny + 3), z(nx + 2, ny + 3), z(nx + 2, ny + 3)
        self.ct = np.zeros((nx + 2, ny + 2), dtype=np.int32)
        self.nu, self.nv, self.np_ = (nx + 1) * ny, nx * (ny + 1), nx * ny
        self.Au, self.Mu = SparseMat(self.nu), SparseMat(self.nu)
        self.Av, self.Mv = SparseMat(self.nv), SparseMat(self.nv)
        self.Ap, self.Mp = SparseMat(self.np_), SparseMat(self.np_)
        self.bu, self.xu = z(self.nu), z(self.nu)
        self.bv, self.xv = z(self.nv), z(self.nv)
        self.bp, self.xp = z(self.np_), z(self.np_)
        self.last_iters = {'u': 0, 'v': 0, 'p': 0}
        self.last_status = {'u': 0, 'v': 0, 'p': 0}

    # karman_taichi.py:130-154
    def init(self):
        nx, ny = self.nx, self.ny
        for i in range(nx + 2):
            for j in range(ny + 2):
                self.p[i, j] = 100 - 4000.0 * i / nx
        self.u[:] = 0.0; self.u0[:] = 0.0; self.v[:] = 0.0; self.v0[:] = 0.0
        self.ct[:] = 1
        self.ct[1:nx + 1, 1:ny + 1] = -1
        cx, cy, r2 = self.cyl
        for i in range(nx):
            for j in range(ny):
                if ((i - cx) ** 2 + (j - cy) ** 2) < r2:
                    self.ct[i, j] = 1
                    self.u[i, j] = 0; self.u0[i, j] = 0; self.v[i, j] = 0; self.v0[i, j] = 0

    # karman_taichi.py:161-248
    def fill_Au(self):
        nx, ny, dx, dy, dt, rho, mu = self.nx, self.ny, self.dx, self.dy, self.dt, self.rho, self.mu
        ct, u, v, p, u0, Au, Mu, bu = self.ct, self.u, self.v, self.p, self.u0, self.Au, self.Mu, self.bu
        for i in range(1, nx + 2):
            for j in range(1, ny + 1):
                k = (i - 1) * ny + (j - 1)
                if ct[i - 1, j] == 1:
                    Au[k, k] = 1.0
                    Au[k, k + ny] = -1.0
                    bu[k] = 0.0
                elif ct[i, j] == 1:
                    Au[k, k] = 1.0
                    Au[k, k - ny] = -1.0
                    bu[k] = 0.0
                elif (ct[i, j] + ct[i, j - 1]) == 0:
                    Au[k, k + 1] = -mu * dx / dy - rho * 0.5 * (v[i - 1, j + 1] + v[i, j + 1]) * dx
                    Au[k, k - ny] = -mu * dy / dx - rho * 0.5 * (u[i, j] + u[i - 1, j]) * dy
                    Au[k, k + ny] = -mu * dy / dx - (-rho * 0.5 * (u[i, j] + u[i + 1, j]) * dy)
                    Au[k, k] = - Au[k, k + 1] - Au[k, k - ny] - \
                        Au[k, k + ny] + rho * dx * dy / dt + 2 * mu * dx / dy
                    bu[k] = (p[i - 1, j] - p[i, j]) * dy + rho * dx * \
                        dy / dt * u0[i, j]
                elif (ct[i, j] + ct[i, j + 1]) == 0:
                    Au[k, k - 1] = -mu * dx / dy - (-rho * 0.5 * (v[i - 1, j] + v[i, j]) * dx)
                    Au[k, k - ny] = -mu * dy / dx - rho * 0.5 * (u[i, j] + u[i - 1, j]) * dy
                    Au[k, k + ny] = -mu * dy / dx - (-rho * 0.5 * (u[i, j] + u[i + 1, j]) * dy)
                    Au[k, k] = -Au[k, k - 1] - Au[k, k - ny] - \
                        Au[k, k + ny] + rho * dx * dy / dt + 2 * mu * dx / dy
                    bu[k] = (p[i - 1, j] - p[i, j]) * dy + rho * dx * \
                        dy / dt * u0[i, j]
                else:
                    Au[k, k - 1] = -mu * dx / dy - (-rho * 0.5 * (v[i - 1, j] + v[i, j]) * dx)
                    Au[k, k + 1] = -mu * dx / dy - rho * 0.5 * (v[i - 1, j + 1] + v[i, j + 1]) * dx
                    Au[k, k - ny] = -mu * dy / dx - rho * 0.5 * (u[i, j] + u[i - 1, j]) * dy
                    Au[k, k + ny] = -mu * dy / dx - (-rho * 0.5 * (u[i, j] + u[i + 1, j]) * dy)
                    Au[k, k] = -Au[k, k - 1] - Au[k, k + 1] - Au[k, k - ny] - \
                        Au[k, k + ny] + rho * dx * dy / dt
                    bu[k] = (p[i - 1, j] - p[i, j]) * dy + rho * dx * \
                        dy / dt * u0[i, j]
                Mu[k, k] = Au[k, k]
        for i in range(2, nx + 1):
            for j in range(2, ny):
                k = (i - 1) * ny + (j - 1)
                if (ct[i - 1, j] + ct[i, j]) == 0 or (ct[i - 1, j] + ct[i, j]) == 2:
                    for nb in [k - ny, k - 1, k + 1, k + ny]:
                        Au[k, nb] = 0.0
                    Au[k, k] = 1.0
                    bu[k] = 0.0
                elif (ct[i, j] + ct[i, j - 1]) == 0 or (ct[i - 1, j] + ct[i - 1, j - 1]) == 0:
                    Au[k, k + 1] = -mu * dx / dy - rho * 0.5 * (v[i - 1, j + 1] + v[i, j + 1]) * dx
                    Au[k, k - ny] = -mu * dy / dx - rho * 0.5 * (u[i, j] + u[i - 1, j]) * dy
                    Au[k, k + ny] = -mu * dy / dx - (-rho * 0.5 * (u[i, j] + u[i + 1, j]) * dy)
                    Au[k, k] = - Au[k, k + 1] - Au[k, k - ny] - \
                        Au[k, k + ny] + rho * dx * dy / dt + 2 * mu * dx / dy
                    bu[k] = (p[i - 1, j] - p[i, j]) * dy + rho * dx * \
                        dy / dt * u0[i, j]
                elif (ct[i, j] + ct[i, j - 1]) == 0 or (ct[i - 1, j] + ct[i - 1, j - 1]) == 0:   # dead: same test (:241)
                    pass

    # karman_taichi.py:250-294
    def fill_Av(self):
        nx, ny, dx, dy, dt, rho, mu = self.nx, self.ny, self.dx, self.dy, self.dt, self.rho, self.mu
        ct, u, v, p, v0, Av, Mv, bv = self.ct, self.u, self.v, self.p, self.v0, self.Av, self.Mv, self.bv
        for i in range(1, nx + 1):
            for j in range(1, ny + 2):
                k = (i - 1) * (ny + 1) + (j - 1)
                if (ct[i, j] + ct[i, j - 1]) == 0 or (ct[i, j] + ct[i, j - 1]) == 2:
                    Av[k, k] = 1.0
                    bv[k] = v[i, j]
                elif (ct[i, j] + ct[i - 1, j]) == 0:
                    Av[k, k - 1] = -mu * dx / dy - (-rho * 0.5 * (v[i, j - 1] + v[i, j]) * dx)
                    Av[k, k + 1] = -mu * dx / dy - (rho * 0.5 * (v[i, j + 1] + v[i, j]) * dx)
                    Av[k, k + ny + 1] = -mu * dy / dx - (-rho * 0.5 * (u[i + 1, j - 1] + u[i + 1, j]) * dy)
                    Av[k, k] = -Av[k, k - 1] - Av[k, k + 1] - \
                        Av[k, k + ny + 1] + rho * dx * dy / dt + 2 * mu * dy / dx
                    bv[k] = (p[i, j] - p[i, j - 1]) * dx + \
                        rho * dx * dy / dt * v0[i, j]
                elif (ct[i, j] + ct[i + 1, j]) == 0:
                    Av[k, k - 1] = -mu * dx / dy - (-rho * 0.5 * (v[i, j - 1] + v[i, j]) * dx)
                    Av[k, k + 1] = -mu * dx / dy - (rho * 0.5 * (v[i, j + 1] + v[i, j]) * dx)
                    Av[k, k - ny - 1] = -mu * dy / dx - (rho * 0.5 * (u[i, j] + u[i, j - 1]) * dy)
                    Av[k, k] = -Av[k, k - 1] - Av[k, k + 1] - Av[k, k - ny - 1] \
                        + rho * dx * dy / dt + 2 * mu * dy / dx
                    bv[k] = (p[i, j] - p[i, j - 1]) * dx + \
                        rho * dx * dy / dt * v0[i, j]
                else:
                    Av[k, k - 1] = -mu * dx / dy - (-rho * 0.5 * (v[i, j - 1] + v[i, j]) * dx)
                    Av[k, k + 1] = -mu * dx / dy - (rho * 0.5 * (v[i, j + 1] + v[i, j]) * dx)
                    Av[k, k - ny - 1] = -mu * dy / dx - (rho * 0.5 * (u[i, j] + u[i, j - 1]) * dy)
                    Av[k, k + ny + 1] = -mu * dy / dx - (-rho * 0.5 * (u[i + 1, j - 1] + u[i + 1, j]) * dy)
                    Av[k, k] = -Av[k, k - 1] - Av[k, k + 1] - Av[k, k - ny - 1] - \
                        Av[k, k + ny + 1] + rho * dx * dy / dt
                    bv[k] = (p[i, j] - p[i, j - 1]) * dx + \
                        rho * dx * dy / dt * v0[i, j]
        for i in range(1, nx + 1):
            for j in range(1, ny + 2):
                k = (i - 1) * (ny + 1) + (j - 1)
                Mv[k, k] = Av[k, k]

    # karman_taichi.py:296-339
    def fill_Ap(self):
        nx, ny, dx, dy, rho = self.nx, self.ny, self.dx, self.dy, self.rho
        ct, u, v, Au, Av, Ap, Mp, bp = self.ct, self.u, self.v, self.Au, self.Av, self.Ap, self.Mp, self.bp
        with np.errstate(divide='ignore', invalid='ignore'):
            for i in range(1, nx + 1):
                for j in range(1, ny + 1):
                    k = (i - 1) * ny + (j - 1)
                    bp[k] = rho * (u[i, j] - u[i + 1, j]) * dy + rho * (v[i, j + 1] - v[i, j]) * dx
                    vk = (i - 1) * (ny + 1) + (j - 1)
                    Ap[k, k - 1] = np.float64(-rho * dx * dx) / np.float64(Av[vk, vk])
                    Ap[k, k + 1] = np.float64(-rho * dx * dx) / np.float64(Av[vk + 1, vk + 1])
                    uk = k
                    Ap[k, k - ny] = np.float64(-rho * dy * dy) / np.float64(Au[uk, uk])
                    Ap[k, k + ny] = np.float64(-rho * dy * dy) / np.float64(Au[uk + ny, uk + ny])
                    if (ct[i, j] + ct[i, j - 1]) == 0:
                        Ap[k, k - 1] = 0
                    if (ct[i, j] + ct[i, j + 1]) == 0:
                        Ap[k, k + 1] = 0
                    if (ct[i, j] + ct[i - 1, j]) == 0:
                        Ap[k, k - ny] = 0
                    if (ct[i, j] + ct[i + 1, j]) == 0:
                        Ap[k, k + ny] = 0
                    Ap[k, k] = -Ap[k, k - 1] - Ap[k, k + 1] - Ap[k, k - ny] - Ap[k, k + ny]
                    Mp[k, k] = Ap[k, k]
        n = nx * ny
        for i in range(ny):
            for j in range(n):
                if i == j:
                    Ap[i, j] = 1.0
                    Mp[i, j] = 1.0
                    bp[j] = 0.0
                    Ap[n - i, n - j] = 1.0
                    Mp[n - i, n - j] = 1.0
                    if 0 <= n - j < n:
                        bp[n - j] = 0.0
                elif (i, j) in Ap.d:            # only materialised entries matter (the rest already read as 0)
                    Ap[i, j] = 0.0
                    Mp[i, j] = 0.0

    # karman_taichi.py:341-456 -- the 7-offset products through CSR
    def bicgstab(self, A, b, x, M, ny, n, eps):
        A7, missed = A.to_csr7(ny)
        assert not missed, f'non-zero entries outside the 7 offsets of karman_taichi.py:365: {missed[:3]}'
        Md = M.diag()
        one = np.float64(1.0)
        with np.errstate(all='ignore'):
            Ax = A7 @ x
            r = b - Ax
            r_tld = r.copy()
            residual_init = np.float64(krylov._sum(r * r))
            omega = alpha = beta = rho_1 = one
            p = np.zeros(n); Ap = np.zeros(n)
            iters, status = 0, 1
            for steps in range(100 * n):
                rho = np.float64(krylov._sum(r * r_tld))
                if rho == 0.0:
                    status = 2
                    break
                if steps == 0:
                    p = r.copy()
                else:
                    beta = (rho / rho_1) * (alpha / omega)
                    p = r + beta * (p - omega * Ap)
                p_hat = 1 / Md * p
                Ap = A7 @ p_hat
                alpha_lower = np.float64(krylov._sum(r_tld * Ap))
                alpha = rho / alpha_lower
                s = r - alpha * Ap
                s_hat = 1 / Md * s
                t = A7 @ s_hat
                omega = np.float64(krylov._sum(t * s)) / np.float64(krylov._sum(t * t))
                x += alpha * p_hat + omega * s_hat
                r = s - omega * t
                residual = np.float64(krylov._sum(r * r))
                iters = steps + 1
                if np.sqrt(residual / residual_init) < eps:
                    status = 0
                    break
                if omega == 0.0:
                    status = 6
                    break
                if not np.isfinite(residual):
                    status = 5
                    break
                rho_1 = rho
        return iters, status

    # karman_taichi.py:458-504
    def xu_back(self):
        nx, ny = self.nx, self.ny
        xu = self.xu.reshape(nx + 1, ny)
        U = self.u[1:nx + 2, 1:ny + 1]
        U[:] = U + self.velo_rel * (xu - U)

    def xv_back(self):
        nx, ny = self.nx, self.ny
        xv = self.xv.reshape(nx, ny + 1)
        V = self.v[1:nx + 1, 1:ny + 2]
        V[:] = V + self.velo_rel * (xv - V)

    def xp_back(self):
        nx, ny = self.nx, self.ny
        self.pcor[1:nx + 1, 1:ny + 1] = self.xp.reshape(nx, ny)

    def solve_momentum_bicgstab(self, eps, max_iterations):
        for _ in range(max_iterations):
            self.fill_Au()
            self.fill_Av()
            self.last_iters['u'], self.last_status['u'] = self.bicgstab(self.Au, self.bu, self.xu, self.Mu, self.ny, self.nu, eps)
            self.last_iters['v'], self.last_status['v'] = self.bicgstab(self.Av, self.bv, self.xv, self.Mv, self.ny, self.nv, eps)
            self.xu_back()
            self.xv_back()

    def solve_pcorrection_bicgstab(self, eps):
        self.fill_Ap()
        self.last_iters['p'], self.last_status['p'] = self.bicgstab(self.Ap, self.bp, self.xp, self.Mp, self.ny, self.np_, eps)
        self.xp_back()

    # karman_taichi.py:506-542
    def puv_correction(self):
        nx, ny, dx, dy = self.nx, self.ny, self.dx, self.dy
        ct, u, v, p, pcor = self.ct, self.u, self.v, self.p, self.pcor
        for i in range(1, nx + 2):
            for j in range(1, ny + 1):
                k = (i - 1) * ny + (j - 1)
                if (ct[i - 1, j] + ct[i, j]) == 0 or (ct[i - 1, j] + ct[i, j]) == 2:
                    pass
                else:
                    self.ucor[i, j] = (pcor[i - 1, j] - pcor[i, j]) * dy / self.Au[k, k]
                    u[i, j] = u[i, j] + self.ucor[i, j]
        for i in range(1, nx + 1):
            for j in range(1, ny + 2):
                k = (i - 1) * (ny + 1) + (j - 1)
                if (ct[i, j] + ct[i, j - 1]) == 0 or (ct[i, j] + ct[i, j - 1]) == 2:
                    pass
                else:
                    self.vcor[i, j] = (pcor[i, j] - pcor[i, j - 1]) * dx / self.Av[k, k]
                    v[i, j] = v[i, j] + self.vcor[i, j]
        pcor_max = 0.0
        for i in range(nx + 2):
            for j in range(ny + 2):
                if ct[i, j] != 1:
                    p[i, j] = p[i, j] + self.p_rel * pcor[i, j]
                if abs(pcor[i, j]) > pcor_max:
                    pcor_max = abs(pcor[i, j])
        return pcor_max

    # karman_taichi.py:585-614
    def correct_conserv(self):
        nx, ny, dx, dy = self.nx, self.ny, self.dx, self.dy
        u, v = self.u, self.v
        mdot_inlet = float(krylov._sum(u[1, 1:ny + 1]))
        mdot_outlet = float(krylov._sum(u[nx + 1, 1:ny + 1]))
        with np.errstate(all='ignore'):
            coef = np.float64(mdot_inlet) / np.float64(mdot_outlet)
        u[nx + 1, 1:ny + 1] = coef * u[nx + 1, 1:ny + 1]
        self.udiv[1:nx + 1, 1:ny + 1] = (u[1:nx + 1, 1:ny + 1] - u[2:nx + 2, 1:ny + 1]) * dy + \
            (v[1:nx + 1, 2:ny + 2] - v[1:nx + 1, 1:ny + 1]) * dx

    # karman_taichi.py:616-622
    def time_forward(self):
        self.u0[:] = self.u
        self.v0[:] = self.v

    def sub_iteration(self, eps=1e-8):
        """One pass of the inner loop body, ``karman_taichi.py:633-652``; returns max |pcor|."""
        self.solve_momentum_bicgstab(eps, self.momentum_iterations)
        self.correct_conserv()
        self.solve_pcorrection_bicgstab(eps)
        return self.puv_correction()

    def time_step(self, max_sub=500, eps=1e-8):
        """``karman_taichi.py:628-660``; returns the number of sub-iterations run."""
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
