"""``SlabKarmanSolver`` -- the channel + cylinder SIMPLE loop of ``karman_taichi.py`` cut into x-slabs, one per GPU
(BASELINE config 4 on 2-8 B200; SURVEY.md 8e: the reference is a single address space and has no counterpart).

Every rank holds the global rows ``[g0, g1)`` of every field: its block of cell rows plus one halo row on each interior
side (the u-type planes, which have one row more than the cell-centred ones, are cut at the same rows; the extra row goes
to the last rank).  The kernels are the single-GPU ones with a row window (``kt_channel_params.row0 / row_lo / row_hi``):
all boundary tests of the reference stay on the GLOBAL row index, so assembly, relaxation and corrections are bit-identical
to the one-GPU run.  Between them the boundary rows of ``u, v, p, p'`` and of the ``ap_u`` plane travel over NVLink
(``kt_p2p_exchange``), the inlet / outlet column sums of ``correct_conserv`` and ``max |p'|`` are all-reduced in rank order
(``kt_p2p_allreduce``), and the three linear systems are solved by the peer-memory BiCGSTAB (``P2PSlabSolver``: Jacobi
plane, warm start, ``omega == 0`` exit -- ``karman_taichi.py:342-456``).

One process per GPU under ``torch.distributed`` (``peers='ipc'``); ``peers='local'`` runs all ranks inside one process
(a thread and a stream each) -- how the comm kernels are tested on a one-GPU box.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .cgsolver import _Handle, _p
from .field import Field
from .slab import P2PSlabSolver, SlabPartition


class SlabKarmanSolver:
    def __init__(self, rank, nranks, nx=320, ny=64, lx=0.5, ly=0.1, cyl=(64, 32, 144), dt=0.00001, momentum_iterations=1,
                 dtype=np.float64, device=None, group=None, peers='ipc'):
        self.rank, self.nranks = int(rank), int(nranks)
        self.nx, self.ny, self.lx, self.ly = nx, ny, lx, ly
        self.dx, self.dy = lx / nx, ly / ny
        self.rho, self.mu, self.dt = 1.0, 0.01, dt
        self.velo_rel, self.p_rel = 0.8, 0.2
        self.cyl = tuple(int(c) for c in cyl)
        self.momentum_iterations = momentum_iterations
        self.real = np.dtype(dtype)
        # cell rows 0 .. nx+1 cut into nranks blocks; the u-type planes (rows 0 .. nx+2) are cut at the same rows
        self.part_p = SlabPartition(nx + 2, nranks)
        self.part_u = SlabPartition.from_bounds(self.part_p.bounds[:-1] + [nx + 3])
        self.g0, g1p = self.part_p.stored(rank)
        g0u, g1u = self.part_u.stored(rank)
        assert g0u == self.g0
        self.rows_p, self.rows_u = g1p - self.g0, g1u - self.g0
        self.own_lo, self.own_hi = self.part_u.owned(rank)          # window of the element-wise kernels (u rows reach furthest)
        mk = lambda rows, cols, *k: Field((rows, cols) + k, self.real, device)
        ru, rp = self.rows_u, self.rows_p
        self.p, self.pcor, self.udiv = mk(rp, ny + 2), mk(rp, ny + 2), mk(rp, ny + 2)
        self.u, self.u0 = mk(ru, ny + 2), mk(ru, ny + 2)
        self.v, self.v0 = mk(rp, ny + 3), mk(rp, ny + 3)
        self.ct = torch.zeros((rp, self.p.pitch), dtype=torch.int32, device=self.p.device)
        self.coef_u, self.b_u, self.M_u, self.Minv_u, self.x_u = mk(ru, ny + 2, 5), mk(ru, ny + 2), mk(ru, ny + 2), mk(ru, ny + 2), mk(ru, ny + 2)
        self.coef_v, self.b_v, self.M_v, self.Minv_v, self.x_v = mk(rp, ny + 3, 5), mk(rp, ny + 3), mk(rp, ny + 3), mk(rp, ny + 3), mk(rp, ny + 3)
        self.coef_p, self.b_p, self.M_p, self.Minv_p = mk(rp, ny + 2, 5), mk(rp, ny + 2), mk(rp, ny + 2), mk(rp, ny + 2)
        self.x_p = self.pcor                                   # xp_back is a copy of xp into pcor (:473-476): alias
        self._hd = _Handle(self.p)
        self._lib = self._hd.lib
        self.sums = torch.zeros(2, dtype=torch.float64, device=self.p.device)
        mkslv = lambda coef, b, x, minv, part, n: self._jacobi_solver(coef, b, x, minv, part, n, group, peers)
        self.u_solver = mkslv(self.coef_u, self.b_u, self.x_u, self.Minv_u, self.part_u, (nx + 1) * ny)
        self.v_solver = mkslv(self.coef_v, self.b_v, self.x_v, self.Minv_v, self.part_p, nx * (ny + 1))
        self.p_solver = mkslv(self.coef_p, self.b_p, self.x_p, self.Minv_p, self.part_p, nx * ny)
        self.last_iters = {'u': 0, 'v': 0, 'p': 0}
        self.last_status = {'u': 0, 'v': 0, 'p': 0}
        self.max_steps = None                                  # cap of every Krylov solve (bench.py: config 4)

    def _jacobi_solver(self, coef, b, x, minv, part, n, group, peers):
        s = P2PSlabSolver(coef, b, x, part, self.rank, group=group, peers=peers)
        s.Minv, s.warm_start, s.omega_zero_break = minv, True, True
        s.default_max = 100 * n                                # karman_taichi.py:382
        s._ws()                                                # the r~ plane of the warm start exists before the first solve
        return s

    @staticmethod
    def connect_local(solvers):
        for name in ('u_solver', 'v_solver', 'p_solver'):
            P2PSlabSolver.connect_local([getattr(s, name) for s in solvers])

    # -- plumbing -------------------------------------------------------------------------------------------------------
    @property
    def _h(self):
        self._hd.use_current_stream()
        return self._hd.h

    def _params(self, lo=None, hi=None):
        P = _lib.ChannelParams()
        P.nx, P.ny = self.nx, self.ny
        P.dx, P.dy, P.dt, P.rho, P.mu = self.dx, self.dy, float(self.dt), self.rho, self.mu
        P.velo_rel, P.p_rel = self.velo_rel, self.p_rel
        P.cyl_cx, P.cyl_cy, P.cyl_r2 = self.cyl
        P.pitch_u, P.pitch_v, P.pitch_p = self.u.pitch, self.v.pitch, self.p.pitch
        P.row0 = self.g0
        P.row_lo, P.row_hi = (self.own_lo if lo is None else lo), (self.own_hi if hi is None else hi)
        P.rows_u, P.rows_p = self.rows_u, self.rows_p
        return P

    def _ct(self):
        return C.c_void_p(self.ct.data_ptr())

    def owned(self, field):
        """The rows of a local field this rank owns, as a NumPy array (tests / gathering)."""
        u_type = any(field is f for f in (self.u, self.u0, self.x_u, self.b_u, self.M_u, self.Minv_u, self.coef_u))
        l0, l1 = (self.part_u if u_type else self.part_p).local_owned(self.rank)
        return field.to_numpy()[l0:l1]

    # -- reference functions ------------------------------------------------------------------------------------------------
    def init(self):
        """``karman_taichi.py:130-154`` on every stored row (owned + halo: the formula depends on the global index only)"""
        for f in (self.pcor, self.x_u, self.x_v):
            f.data.zero_()
        P = self._params(self.g0, self.g0 + self.rows_u)
        _lib.check(self._lib.kt_channel_init(self._h, C.byref(P), _p(self.p), _p(self.u), _p(self.u0), _p(self.v), _p(self.v0), self._ct()))

    def fill_Au(self):
        """``karman_taichi.py:161-294`` (fill_Au + fill_Av) on the owned rows; reads the halo rows of u, v, p, ct"""
        P = self._params()
        _lib.check(self._lib.kt_channel_assemble_uv(
            self._h, C.byref(P), self._ct(), _p(self.u), _p(self.v), _p(self.p), _p(self.u0), _p(self.v0),
            _p(self.coef_u), _p(self.b_u), _p(self.M_u), _p(self.Minv_u), _p(self.coef_v), _p(self.b_v), _p(self.M_v), _p(self.Minv_v)))
        # the solver's operand is M^-1 p: the product A (M^-1 p) reads M^-1 on the halo rows too (karman_taichi.py:399-406)
        self.u_solver.exchange(self.Minv_u.ptr)
        self.v_solver.exchange(self.Minv_v.ptr)

    fill_Av = fill_Au

    def fill_Ap(self):
        """``karman_taichi.py:296-339``; needs ap_u of the face above the slab: the centre plane of coef_u is exchanged first"""
        self.u_solver.exchange(self.coef_u.ptr)               # plane 0 (ap) of the five
        P = self._params()
        _lib.check(self._lib.kt_channel_assemble_p(self._h, C.byref(P), self._ct(), _p(self.u), _p(self.v), _p(self.coef_u),
                                                   _p(self.coef_v), _p(self.coef_p), _p(self.b_p), _p(self.M_p), _p(self.Minv_p)))
        self.p_solver.exchange(self.Minv_p.ptr)

    def bicgstab(self, which, eps):
        """``karman_taichi.py:341-456`` for system 'u', 'v' or 'p', over all ranks"""
        s = {'u': self.u_solver, 'v': self.v_solver, 'p': self.p_solver}[which]
        cap = s.default_max if self.max_steps is None else self.max_steps
        s.solve(eps=eps, max_iter=cap)
        self.last_iters[which], self.last_status[which] = s.iterations, s.status

    def xu_back(self):
        """``karman_taichi.py:458-470`` on the owned rows, then the new boundary rows of u and v go to the neighbours"""
        P = self._params()
        _lib.check(self._lib.kt_channel_relax(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.x_u), _p(self.x_v)))
        self.u_solver.exchange(self.u.ptr)
        self.v_solver.exchange(self.v.ptr)

    xv_back = xu_back

    def solve_momentum_bicgstab(self, eps, max_iterations):
        """``karman_taichi.py:480-494``"""
        for _ in range(max_iterations):
            self.fill_Au()
            self.bicgstab('u', eps)
            self.bicgstab('v', eps)
            self.xu_back()

    def correct_conserv(self):
        """``karman_taichi.py:585-614``: inlet sum on the first rank, outlet sum on the last, added in rank order"""
        P = self._params()
        _lib.check(self._lib.kt_outlet_sums(self._h, C.byref(P), _p(self.u), C.c_void_p(self.sums.data_ptr())))
        sa, sb = self.u_solver.allreduce(0.0, 0.0, 'sum', device_pair=self.sums)
        _lib.check(self._lib.kt_outlet_apply(self._h, C.byref(P), _p(self.u), sa, sb))
        self.u_solver.exchange(self.u.ptr)
        _lib.check(self._lib.kt_channel_udiv(self._h, C.byref(P), _p(self.u), _p(self.v), _p(self.udiv)))
        return sa, sb, sa / sb if sb != 0.0 else float('nan')

    def solve_pcorrection_bicgstab(self, eps):
        """``karman_taichi.py:497-504``"""
        self.fill_Ap()
        self.bicgstab('p', eps)

    def puv_correction(self):
        """``karman_taichi.py:506-542``; returns max |pcor| over all ranks"""
        self.p_solver.exchange(self.pcor.ptr)                 # u[i] reads p'[i-1]
        P = self._params()
        out = C.c_double()
        _lib.check(self._lib.kt_channel_correct(self._h, C.byref(P), self._ct(), _p(self.u), _p(self.v), _p(self.p), _p(self.pcor),
                                                _p(self.coef_u), _p(self.coef_v), C.byref(out)))
        mx, _ = self.p_solver.allreduce(out.value, 0.0, 'max')
        self.u_solver.exchange(self.u.ptr)
        self.v_solver.exchange(self.v.ptr)
        self.p_solver.exchange(self.p.ptr)
        return mx

    def time_forward(self):
        """``karman_taichi.py:616-622``"""
        self.u0.copy_from(self.u)
        self.v0.copy_from(self.v)

    def sub_iteration(self, eps=1e-8):
        """main loop body, ``karman_taichi.py:634-653``"""
        self.solve_momentum_bicgstab(eps, self.momentum_iterations)
        self.correct_conserv()
        self.solve_pcorrection_bicgstab(eps)
        return self.puv_correction()

    def time_step(self, max_sub=500, eps=1e-8):
        pcor_max, n = 0.0, 0
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
