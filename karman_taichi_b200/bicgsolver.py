"""``BICGSolver`` -- host mirror of the reference class (``bicgsolver.py:6-135``) over libkarman_b200.

Unpreconditioned BiCGSTAB on the 5-point stencil arrays, x0 = 0, exactly the reference's recurrence;
``solve()`` is one device-resident call (``kt_bicgstab_solve``).  The identity copies of the reference
(``p_hat = p``, ``s_hat = s``, ``t = Ashat``: ``bicgsolver.py:48-58,114``) are aliases here, and ``r_tld`` aliases
``b`` (it is a copy of ``b`` that is never modified: ``:25-26``), which saves 4 N words per iteration and four
planes of HBM.  With ``Minv`` / ``warm_start`` / ``omega_zero_break`` it runs the monolith's Jacobi variant
(``karman_taichi.py:342-456``).
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from .cgsolver import CGSolver, _p
from .field import Field, ScalarField


class BICGSolver(CGSolver):
    def __init__(self, coef, b, x):
        super().__init__(coef, b, x)
        mk = lambda: Field((self.NX, self.NY), self.real, coef.device)
        self.r_tld = self.b          # alias (see module docstring)
        self.p_hat = self.p          # bicgsolver.py:48-52 identity
        self._s = None               # allocated on first use: solve() does not store s unless store_s is set
        self.t = mk()
        self.Ashat = self.t          # bicgsolver.py:114 copy(Ashat, t)
        self.omega = ScalarField(0.0)
        self.rho = ScalarField(0.0)
        self.rho_1 = ScalarField(0.0)
        # Jacobi variant (karman_taichi.py:342-456); None / False = class version
        self.Minv = None
        self.warm_start = False
        self.omega_zero_break = False
        self._r_tld_own = None
        self.max_steps = None
        # fused K1+K2 (27 N words / iteration, two extra planes).  Off by default: measured on B200 the 12-stream fused
        # kernel streams at 6.0 TB/s vs 6.7-7.0 TB/s for K1 and K2 apart -- 9.29 vs 9.00 ms / iteration at 16384^2
        # (profiles/r01_grid_sweep.md).  The persistent small-grid solver uses it (one grid barrier less).
        self.fuse_p_update = False
        self._p2 = self._Ap2 = None
        # s = r - alpha v is recomputed by the x/r kernel instead of being written by the s,t kernel and read back: same
        # 28 N words per iteration, one write stream less (-1.7 % time at 16384^2) and one plane of HBM saved.
        self.store_s = False
        self.report_rho_zero = True   # bicgsolver.py:97 prints; the monolith only does with output=1 (karman_taichi.py:387)

    @property
    def s(self):
        """``bicgsolver.py:11``"""
        if self._s is None:
            self._s = Field((self.NX, self.NY), self.real, self.coef.device)
        return self._s

    @property
    def s_hat(self):
        """``bicgsolver.py:12`` -- identity copy of s (``:54-58``): an alias"""
        return self.s

    # -- reference kernels ---------------------------------------------------------------------
    def init(self):
        """``bicgsolver.py:21-36``"""
        L, h = self._lib, self._h
        _lib.check(L.kt_copy(h, _p(self.r), _p(self.b)))
        for f in (self.Ap, self.p, self.x, self.t):
            _lib.check(L.kt_fill(h, _p(f), 0.0))
        self.omega[None] = 1.0
        self.alpha[None] = 1.0
        self.beta[None] = 1.0
        self.rho_1[None] = 1.0
        self.rho[None] = 0.0

    def copy(self, orig, dest):
        """``bicgsolver.py:38-41``"""
        if orig is not dest:
            _lib.check(self._lib.kt_copy(self._h, _p(dest), _p(orig)))

    def update_p(self):
        """``bicgsolver.py:43-46``"""
        _lib.check(self._lib.kt_bicg_update_p(self._h, _p(self.p), _p(self.r), _p(self.Ap), self.beta[None], self.omega[None]))

    def update_phat(self):
        """``bicgsolver.py:48-52``: identity (p_hat aliases p)"""

    def update_shat(self):
        """``bicgsolver.py:54-58``: identity (s_hat aliases s)"""

    def update_s(self):
        """``bicgsolver.py:60-63``"""
        _lib.check(self._lib.kt_xmay(self._h, _p(self.s), _p(self.r), self.alpha[None], _p(self.Ap)))

    def compute_As(self):
        """``bicgsolver.py:65-69``"""
        _lib.check(self._lib.kt_spmv(self._h, _p(self.coef), _p(self.s_hat), _p(self.Ashat)))

    def compute_Ap(self):
        """``bicgsolver.py:72-76``"""
        _lib.check(self._lib.kt_spmv(self._h, _p(self.coef), _p(self.p_hat), _p(self.Ap)))

    def update_x(self):
        """``bicgsolver.py:78-81``"""
        _lib.check(self._lib.kt_bicg_update_x(self._h, _p(self.x), _p(self.p_hat), _p(self.s_hat), self.alpha[None], self.omega[None]))

    def update_r(self):
        """``bicgsolver.py:83-86``"""
        _lib.check(self._lib.kt_xmay(self._h, _p(self.r), _p(self.s), self.omega[None], _p(self.t)))

    # -- solve ---------------------------------------------------------------------------------
    def solve(self, eps=1e-8, quiet=True):
        """``bicgsolver.py:88-135``"""
        flags = 0
        r_tld = None
        if self.warm_start:
            flags |= _lib.FLAG_WARM_START
            if self._r_tld_own is None:
                self._r_tld_own = Field((self.NX, self.NY), self.real, self.coef.device)
            r_tld = self._r_tld_own.ptr
        if self.omega_zero_break:
            flags |= _lib.FLAG_OMEGA_ZERO_BREAK
        p2 = ap2 = None
        if self.fuse_p_update:
            if self._p2 is None:
                self._p2 = Field((self.NX, self.NY), self.real, self.coef.device)
                self._Ap2 = Field((self.NX, self.NY), self.real, self.coef.device)
            p2, ap2 = self._p2.ptr, self._Ap2.ptr
        ws = self._ws(r_tld=r_tld, s=(self.s.ptr if self.store_s else None), t=self.t.ptr, p2=p2, Ap2=ap2)
        iters, rel, status = C.c_long(), C.c_double(), C.c_int()
        steps = self.steps if self.max_steps is None else self.max_steps
        _lib.check(self._lib.kt_set_cluster(self._h, {True: 1, False: 0, 'coop': 2, 'cluster': 3}[self.cluster]))
        _lib.check(self._lib.kt_bicgstab_solve(self._h, _p(self.coef), _p(self.b), _p(self.x), _p(self.Minv), C.byref(ws),
                                               float(eps), steps, flags, C.byref(iters), C.byref(rel), C.byref(status)))
        hist = self._finish(iters, rel, status, quiet, None, want_history=not quiet)   # history is commented out in the reference (:93,125)
        st = status.value
        if not quiet:
            print('>>> Initial residual =', hist[0] if len(hist) else float('nan'))
            for i in range(1, min(self.iterations, len(hist) - 1) + 1):
                print('>>> Iter =', i, ' Residual =', hist[i])
        if st == _lib.STATUS_RHO_ZERO:
            if self.report_rho_zero:
                print('>>> BICG failed at first place...')           # bicgsolver.py:97
        elif st in (_lib.STATUS_ALPHA_DIV_ZERO, _lib.STATUS_OMEGA_DIV_ZERO):
            raise ZeroDivisionError('float division by zero')    # bicgsolver.py:102,108,118
        elif st == _lib.STATUS_CONVERGED and not quiet:
            print('>>> BICG Converged...')
