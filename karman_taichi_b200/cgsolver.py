"""``CGSolver`` -- host mirror of the reference class (``cgsolver.py:5-131``) over libkarman_b200.

Same constructor, attributes and methods as the reference; every method that was a ``@ti.kernel``
is one C-ABI call (``include/karman_b200.h``), and ``solve()`` is a single device-resident call
(``kt_cg_solve``): the per-iteration host round trips of ``cgsolver.py:92-114`` are gone, alpha/beta and
the convergence test live on the device.  Importing this module runs nothing.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .field import Field, ScalarField

HISTORY_CAP = 1 << 16


class _Handle:
    """RAII wrapper of ``kt_handle_t`` for one (device, NX, NY, pitch, dtype) geometry."""

    def __init__(self, like: Field):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        self.device = like.device
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.kt_create(self.device.index or 0, like.NX, like.NY, like.pitch, like.dtype_code,
                                      C.c_void_p(stream), C.byref(self.h)))

    def use_current_stream(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.kt_set_stream(self.h, C.c_void_p(stream)))

    def launch_count(self):
        return int(self.lib.kt_launch_count(self.h))

    def __del__(self):
        try:
            if self.h:
                self.lib.kt_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


def _p(f):
    return C.c_void_p(f.ptr if f is not None else None)


class CGSolver:
    def __init__(self, coef, b, x):
        if not (isinstance(coef, Field) and coef.is_coef):
            raise TypeError('coef must be a Field of shape (NX, NY, 5)')
        self.coef = coef
        self.b = b
        self.x = x
        self.NX = coef.shape[0]
        self.NY = coef.shape[1]
        if b.shape != (self.NX, self.NY) or x.shape != (self.NX, self.NY):
            raise ValueError('b and x must have shape (NX, NY) of coef')
        self.real = coef.np_dtype
        self.N_ext = 0
        self.steps = self.NX * self.NY      # cgsolver.py:14
        self._history = []                  # cgsolver.py:15 (strings, or raw residual arrays formatted on first read)
        mk = lambda: Field((self.NX, self.NY), self.real, coef.device)
        self.r = mk()
        self.p = mk()
        self.Ap = mk()
        self._Ax = None                     # only compute_Ax() needs it: allocated on first use
        self.alpha = ScalarField(0.0)
        self.beta = ScalarField(0.0)
        self._hd = _Handle(b)
        self._lib = self._hd.lib
        self._hist_dev = None
        # resident solvers: True = automatic (one thread-block cluster for small grids, one persistent cooperative kernel
        # for mid-size ones: kt_cluster.cu / kt_coop.cu), False = streaming kernels only, 'cluster' / 'coop' = only that one
        self.cluster = True
        self.iterations = 0                 # bookkeeping (not in the reference)
        self.status = 0
        self.rel_res = float('nan')

    # -- plumbing ---------------------------------------------------------------------------
    @property
    def history(self):
        """``cgsolver.py:15,88,111``: list of ``f'{residual:e}\\n'`` strings.  Residual arrays coming back from the device
        are kept raw and formatted here on first access (a SIMPLE run appends hundreds of entries per outer iteration)."""
        if any(not isinstance(h, str) for h in self._history):
            flat = []
            for h in self._history:
                if isinstance(h, str):
                    flat.append(h)
                else:
                    flat.extend(f'{v:e}\n' for v in h)
            self._history = flat
        return self._history

    @history.setter
    def history(self, value):
        self._history = list(value)

    @property
    def Ax(self):
        if self._Ax is None:
            self._Ax = Field((self.NX, self.NY), self.real, self.coef.device)
        return self._Ax

    @property
    def _h(self):
        self._hd.use_current_stream()
        return self._hd.h

    def _history_buffer(self):
        cap = min(self.steps + 1, HISTORY_CAP)
        if self._hist_dev is None or self._hist_dev.numel() != cap:
            self._hist_dev = torch.zeros(cap, dtype=torch.float64, device=self.coef.device)
        return self._hist_dev

    def _ws(self, **extra):
        hist = self._history_buffer()
        ws = _lib.KrylovWs(r=self.r.ptr, p=self.p.ptr, Ap=self.Ap.ptr, r_tld=None, s=None, t=None,
                           history=hist.data_ptr(), history_cap=hist.numel(), p2=None, Ap2=None)
        for k, v in extra.items():
            setattr(ws, k, v)
        return ws

    # -- reference kernels, one C-ABI call each ---------------------------------------------
    def init(self):
        """``cgsolver.py:31-39``"""
        L, h = self._lib, self._h
        _lib.check(L.kt_copy(h, _p(self.r), _p(self.b)))
        for f in (self.Ap, self.p, self.x):
            _lib.check(L.kt_fill(h, _p(f), 0.0))
        if self._Ax is not None:
            _lib.check(L.kt_fill(h, _p(self._Ax), 0.0))

    def update_coef(self, coef, b, x):
        """``cgsolver.py:41-47``: copy into the solver's own fields (a self-copy in ``bicgsimp.py:205-208``)."""
        if coef is not self.coef:
            self.coef.copy_from(coef)
        if b is not self.b:
            self.b.copy_from(b)
        if x is not self.x:
            self.x.copy_from(x)

    def reduce(self, p, q):
        """``cgsolver.py:49-54``"""
        out = C.c_double()
        _lib.check(self._lib.kt_dot(self._h, _p(p), _p(q), C.byref(out)))
        return out.value

    def compute_Ap(self):
        """``cgsolver.py:56-60``"""
        _lib.check(self._lib.kt_spmv(self._h, _p(self.coef), _p(self.p), _p(self.Ap)))

    def compute_Ax(self):
        """``cgsolver.py:62-66``"""
        _lib.check(self._lib.kt_spmv(self._h, _p(self.coef), _p(self.x), _p(self.Ax)))

    def update_x(self):
        """``cgsolver.py:68-71``"""
        _lib.check(self._lib.kt_axpy(self._h, _p(self.x), self.alpha[None], _p(self.p)))

    def update_r(self):
        """``cgsolver.py:73-76``"""
        _lib.check(self._lib.kt_axpy(self._h, _p(self.r), -self.alpha[None], _p(self.Ap)))

    def update_p(self):
        """``cgsolver.py:78-81``"""
        _lib.check(self._lib.kt_xpay(self._h, _p(self.p), _p(self.r), self.beta[None], _p(self.p)))

    def compute_residual(self):
        """``cgsolver.py:121-126``: r = b - Ax"""
        _lib.check(self._lib.kt_xmay(self._h, _p(self.r), _p(self.b), 1.0, _p(self.Ax)))

    def check_solution(self):
        """``cgsolver.py:128-131``"""
        return np.sqrt(self.reduce(self.r, self.r))

    # -- solve ---------------------------------------------------------------------------------
    def _finish(self, iters, rel, status, quiet, converged_msg, want_history=True):
        self.iterations, self.rel_res, self.status = iters.value, rel.value, status.value
        if not want_history:
            return None
        return self._hist_dev[:min(self.iterations + 1, self._hist_dev.numel())].cpu().numpy()

    def solve(self, eps=1e-8, quiet=True):
        """``cgsolver.py:83-114``"""
        ws = self._ws()
        iters, rel, status = C.c_long(), C.c_double(), C.c_int()
        _lib.check(self._lib.kt_set_cluster(self._h, {True: 1, False: 0, 'coop': 2, 'cluster': 3}[self.cluster]))
        _lib.check(self._lib.kt_cg_solve(self._h, _p(self.coef), _p(self.b), _p(self.x), C.byref(ws), float(eps),
                                         self.steps, C.byref(iters), C.byref(rel), C.byref(status)))
        hist = self._finish(iters, rel, status, quiet, None)
        if status.value == _lib.STATUS_ALPHA_DIV_ZERO:
            raise ZeroDivisionError('float division by zero')      # cgsolver.py:96 old_rTr / pAp
        n_logged = self.iterations - (1 if status.value == _lib.STATUS_CONVERGED else 0)
        n_logged = min(n_logged, len(hist) - 1)
        self._history.append(hist[:n_logged + 1].copy())         # initial + one entry per non-final iteration (:88,111)
        if not quiet:
            print('>>> Initial residual =', hist[0] if len(hist) else float('nan'))
            for i in range(1, n_logged + 1):
                print(f'>>> Iter = {i:4}, Residual = {hist[i]:e}')
            if status.value == _lib.STATUS_CONVERGED:
                print('>>> Conjugate Gradient method converged.')

    def save_history(self):
        """``cgsolver.py:116-119``"""
        with open('convergence.txt', 'w') as f:
            for line in self.history:
                f.write(line)
