"""``JacobiSolver`` -- the point-Jacobi sweeps of the reference's validation scripts as a solver object on the device.

``pressure_valid.py:220-239`` (``jacobian_solve_p`` / ``jacob_solve_pcorrection_eqn``: sweep until the max-norm of the update
drops below eps) and ``momentum_valid.py:157-170`` (``jacobian_solve_u`` / ``jacob_solve_momentum_eqn``: a fixed number of
under-relaxed sweeps).  Same constructor shape as ``CGSolver(coef, b, x)``; one sweep = ``kt_jacobi_sweep``
(``csrc/kt_post.cu``).  The reference's kernel overwrites ``x`` inside the parallel loop, so its result depends on thread
order (Gauss-Seidel-like when run serially); here every read of a sweep sees the old iterate -- one of the schedules the
reference's loop admits, and a deterministic one.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from .cgsolver import _Handle, _p
from .field import Field


class JacobiSolver:
    def __init__(self, coef, b, x, rows=None, cols=None, relax=1.0):
        if not (isinstance(coef, Field) and coef.is_coef):
            raise TypeError('coef must be a Field of shape (NX, NY, 5)')
        self.coef, self.b, self.x = coef, b, x
        self.NX, self.NY = coef.shape[0], coef.shape[1]
        if b.shape != (self.NX, self.NY) or x.shape != (self.NX, self.NY):
            raise ValueError('b and x must have shape (NX, NY) of coef')
        # the swept range: interior of the plane by default (pressure_valid.py:223 ndrange((1,nx+1),(1,ny+1)))
        self.rows = rows if rows is not None else (1, self.NX - 1)
        self.cols = cols if cols is not None else (1, self.NY - 1)
        self.relax = float(relax)
        self.x_mid = Field((self.NX, self.NY), coef.np_dtype, coef.device)
        self._hd = _Handle(b)
        self._lib = self._hd.lib
        self.sweeps = 0
        self.residual = float('nan')

    def sweep(self):
        """One sweep; returns max |x - x_mid| before the update (``pressure_valid.py:229-232``)."""
        self._hd.use_current_stream()
        out = C.c_double()
        _lib.check(self._lib.kt_jacobi_sweep(self._hd.h, _p(self.coef), _p(self.b), _p(self.x), _p(self.x_mid), self.rows[0], self.rows[1],
                                             self.cols[0], self.cols[1], self.relax, C.byref(out)))
        self.sweeps += 1
        self.residual = out.value
        return out.value

    def solve(self, eps=1e-6, max_sweeps=None, quiet=True, n_iter=None):
        """``jacob_solve_pcorrection_eqn`` (``pressure_valid.py:234-238``): sweep while the residual exceeds eps; or, with
        ``n_iter``, ``jacob_solve_momentum_eqn`` (``momentum_valid.py:172-177``): exactly that many sweeps."""
        self.sweeps = 0
        if n_iter is not None:
            for i in range(n_iter):
                res = self.sweep()
                if not quiet:
                    print('>>> Iter =', i, ', eps =', res)
            return self.sweeps
        res = 1.0
        while res > eps and (max_sweeps is None or self.sweeps < max_sweeps):
            res = self.sweep()
            if not quiet:
                print('>>> Pressure correction residual =', res)
        return self.sweeps
