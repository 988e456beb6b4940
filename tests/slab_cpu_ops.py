"""NumPy stand-in for ``karman_taichi_b200.slab.CudaSlabOps`` built on the oracle (TEST infrastructure).

Lets the x-slab driver (partition, halo exchange, dot all-reduce, scalar algebra order) run on CPU under gloo.
``scalar()`` restates ``kt::scalar_algebra`` (csrc/kt_common.cuh), i.e. the host-scalar lines of
``bicgsolver.py:95-135``.
"""
import math

import numpy as np
import torch

from oracle import krylov


class OracleSlabOps:
    def __init__(self, coef_local, b_local, l0, l1):
        self.coef, self.b = coef_local, b_local
        self.l0, self.l1 = l0, l1
        z = lambda: torch.zeros(b_local.shape, dtype=torch.float64)
        self.x, self.r, self.p, self.v, self.s, self.t = z(), z(), z(), z(), z(), z()
        self.dots = torch.zeros(2, dtype=torch.float64)
        self.sc = {}

    def planes(self):
        return {'p': self.p, 'v': self.v, 'r': self.r}

    def _own(self, a):
        return a[self.l0:self.l1]

    def begin(self, eps, max_iter):
        self.sc = dict(rho=0.0, rho_1=1.0, alpha=1.0, beta=1.0, omega=1.0, rTr0=0.0, rTr=0.0, eps=eps, iter=0,
                       max_iter=max_iter, status=1, done=False)

    def init(self):
        b = torch.from_numpy(self.b)
        self._own(self.r)[:] = self._own(b)
        self._own(self.x)[:] = 0.0
        self.dots[0] = float(np.sum(self._own(self.b) * self._own(self.b)))
        self.dots[1] = 0.0

    def update_p(self):
        if self.sc['done']:
            return
        sc = self.sc
        if sc['iter'] == 0:
            self._own(self.p)[:] = self._own(self.r)
        else:
            self._own(self.p)[:] = self._own(self.r) + sc['beta'] * (self._own(self.p) - sc['omega'] * self._own(self.v))

    def _spmv_owned(self, x):
        # the stencil on the local plane (halo rows present), keep owned rows only
        return krylov.spmv(self.coef, x.numpy())[self.l0:self.l1]

    def v_dot(self):
        if self.sc['done']:
            return
        self._own(self.v)[:] = torch.from_numpy(self._spmv_owned(self.p))
        self.dots[0] = float(np.sum(self._own(self.b) * self._own(self.v).numpy()))

    def st_dots(self):
        if self.sc['done']:
            return
        s_full = self.r - self.sc['alpha'] * self.v          # halo rows recomputed from exchanged r, v
        t = self._spmv_owned(s_full)
        self._own(self.s)[:] = self._own(s_full)
        self._own(self.t)[:] = torch.from_numpy(t)
        so = self._own(self.s).numpy()
        self.dots[0] = float(np.sum(t * so))
        self.dots[1] = float(np.sum(t * t))

    def xr_dots(self):
        if self.sc['done']:
            return
        sc = self.sc
        self._own(self.x)[:] = self._own(self.x) + (sc['alpha'] * self._own(self.p) + sc['omega'] * self._own(self.s))
        self._own(self.r)[:] = self._own(self.s) - sc['omega'] * self._own(self.t)
        rn = self._own(self.r).numpy()
        self.dots[0] = float(np.sum(rn * rn))
        self.dots[1] = float(np.sum(rn * self._own(self.b)))

    # ---- CG (cgsolver.py:83-114) ----
    def cg_init(self):
        b = torch.from_numpy(self.b)
        self._own(self.r)[:] = self._own(b)
        self._own(self.x)[:] = 0.0
        self._own(self.p)[:] = self._own(b)
        self.dots[0] = float(np.sum(self._own(self.b) * self._own(self.b)))
        self.dots[1] = 0.0

    def cg_ap_dot(self):
        if self.sc['done']:
            return
        self._own(self.v)[:] = torch.from_numpy(self._spmv_owned(self.p))
        self.dots[0] = float(np.sum(self._own(self.p).numpy() * self._own(self.v).numpy()))

    def cg_xr_dot(self):
        if self.sc['done']:
            return
        a = self.sc['alpha']
        self._own(self.x)[:] = self._own(self.x) + a * self._own(self.p)
        self._own(self.r)[:] = self._own(self.r) - a * self._own(self.v)
        rn = self._own(self.r).numpy()
        self.dots[0] = float(np.sum(rn * rn))

    def cg_update_p(self):
        if self.sc['done']:
            return
        self._own(self.p)[:] = self._own(self.r) + self.sc['beta'] * self._own(self.p)

    def scalar(self, fin):
        sc = self.sc
        if sc['done']:
            return
        d0, d1 = float(self.dots[0]), float(self.dots[1])
        if fin == 1:      # FIN_BICG_INIT
            sc['rTr0'] = sc['rTr'] = sc['rho'] = d0
            if sc['max_iter'] <= 0:
                sc['status'], sc['done'] = 1, True
            elif d0 == 0.0:
                sc['status'], sc['done'] = 2, True
        elif fin == 2:    # FIN_BICG_ALPHA
            if d0 == 0.0:
                sc['status'], sc['done'] = 3, True
            else:
                sc['alpha'] = sc['rho'] / d0
        elif fin == 3:    # FIN_BICG_OMEGA
            if d1 == 0.0:
                sc['status'], sc['done'] = 4, True
            else:
                sc['omega'] = d0 / d1
        elif fin == 4:    # FIN_BICG_END
            sc['rTr'] = d0
            sc['iter'] += 1
            if math.sqrt(d0 / sc['rTr0']) < sc['eps']:
                sc['status'], sc['done'] = 0, True
                return
            if sc['iter'] >= sc['max_iter']:
                sc['status'], sc['done'] = 1, True
                return
            sc['rho_1'] = sc['rho']
            sc['rho'] = d1
            if d1 == 0.0:
                sc['status'], sc['done'] = 2, True
                return
            sc['beta'] = (sc['rho'] / sc['rho_1']) * (sc['alpha'] / sc['omega'])

        elif fin == 5:    # FIN_CG_INIT
            sc['rTr0'] = sc['rTr'] = sc['old_rTr'] = d0
            if sc['max_iter'] <= 0:
                sc['status'], sc['done'] = 1, True
        elif fin == 6:    # FIN_CG_ALPHA
            if d0 == 0.0:
                sc['status'], sc['done'] = 3, True
            else:
                sc['alpha'] = sc['old_rTr'] / d0
        elif fin == 7:    # FIN_CG_END
            sc['rTr'] = d0
            sc['iter'] += 1
            if math.sqrt(d0) < math.sqrt(sc['rTr0']) * sc['eps']:
                sc['status'], sc['done'] = 0, True
                return
            sc['beta'] = d0 / sc['old_rTr']
            sc['old_rTr'] = d0
            if sc['iter'] >= sc['max_iter']:
                sc['status'], sc['done'] = 1, True

    def poll(self):
        sc = self.sc
        return sc['done'], sc['iter'], sc['status'], math.sqrt(sc['rTr'] / sc['rTr0']) if sc['rTr0'] else float('nan')
