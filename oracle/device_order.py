"""Inner products added in exactly the order libkarman_b200 adds them.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference's ``reduce`` (``cgsolver.py:49-54``) is a parallel
``sum +=`` whose order Taichi leaves unspecified; the library's reductions are deterministic -- the order is a function of
the geometry alone, published by ``kt_reduce_plan`` / ``kt_resident_plan`` (``include/karman_b200.h``; host-only
functions, no device needed).  This module restates that order in NumPy.  With it installed
(``oracle.krylov.set_summation('device', DeviceOrder(...))``) the oracle performs the same floating-point operations in
the same order as the CUDA path (the element-wise work is already bit-identical: ``-fmad=false``), so iterates, residual
histories and iteration counts must agree BIT FOR BIT -- which is what ``tests/test_gpu_exact.py`` asserts.

Orders restated (f64 only):

* streaming kernels (``csrc/kt_common.cuh:block_reduce_finalize``): thread-sequential over its statically assigned
  elements -> ``__shfl_down`` tree over the 32 lanes -> the block's 8 warps in order -> one partial per block -> thread k of
  the finalising block adds partials k, k+256, ... in order -> lane tree -> warps in order.  Element assignment: grid-stride
  over 16-byte vectors for the element-wise kernels, (row-chunk, column-chunk) tiles per warp for the stencil kernels
  (``csrc/kt_krylov.cu:k_stencil``).
* cluster solver (``csrc/kt_cluster.cu:cluster_sum``): thread-sequential -> lane tree -> 16-leaf tree over the warps ->
  16-leaf tree over the CTAs.
* cooperative solver (``csrc/kt_coop.cu:grid_sum``): thread-sequential -> lane tree -> 32-leaf tree over the warps -> lane
  l adds the slots l, l+32, ... of the CTAs in order -> lane tree.
* x-slab solves: each rank's local sum as above over its owned rows, then the ranks in rank order
  (``csrc/kt_krylov.cu:k_p2p_reduce``).

Every "tree" is the same shape: 2^k leaves, level by level ``a[:h] + a[h:]`` (what lane 0 of a ``shfl_down`` tree and every
lane of a ``shfl_xor`` butterfly compute).
"""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np

PLAN = {'dot': 0, 'init': 1, 'alpha': 2, 'omega': 3, 'end': 4, 'end_unit': 5, 'resid': 6, 'cg_init': 1, 'cg_alpha': 7, 'cg_end': 8,
        'alpha_fused': 9}


def tree(a):
    """Halving tree over the last axis (length a power of two)."""
    n = a.shape[-1]
    assert n & (n - 1) == 0
    while n > 1:
        n //= 2
        a = a[..., :n] + a[..., n:2 * n]
    return a[..., 0]


def seq(a):
    """Strictly sequential sum over the last axis, starting from +0.0."""
    acc = np.zeros(a.shape[:-1])
    for k in range(a.shape[-1]):
        acc = acc + a[..., k]
    return acc


def _pad(a, n):
    """Zero-pad the last axis to length n."""
    if a.shape[-1] == n:
        return a
    out = np.zeros(a.shape[:-1] + (n,))
    out[..., :a.shape[-1]] = a
    return out


class _Gather:
    """A fixed assignment of plane elements to (thread, position-in-thread-sequence); -1 = nothing (adds +0.0)."""

    def __init__(self, idx):
        self.idx = idx                      # (..., L) int64 flat indices into the padded plane, -1 = none

    def thread_sums(self, flat):
        ext = np.concatenate([flat, [0.0]])
        return seq(ext[self.idx])


def _lib():
    from karman_taichi_b200 import _lib as L
    return L, L.load()


@contextlib.contextmanager
def installed(order):
    """``with installed(DeviceOrder(...)):`` -- the oracle adds in the device's order inside the block."""
    from . import krylov
    krylov.set_summation('device', order)
    try:
        yield order
    finally:
        krylov.set_summation('pairwise')


class DeviceOrder:
    """``summer(kind, products)`` for one geometry.  ``shape`` = logical (NX, NY) of the planes.

    ``enable``: the solver's ``cluster`` setting (1 automatic, 0 streaming kernels only, 2 cooperative only, 3 cluster
    only).  ``rows`` = (row_lo, row_hi) restricts the streaming plans to a slab's owned rows (``kt_set_rows``)."""

    def __init__(self, shape, is_cg=False, pre=False, warm=False, enable=1, num_sms=148, rows=None, xr_unit=False, fused=False):
        L, lib = _lib()
        self.NX, self.NY = int(shape[0]), int(shape[1])
        self.pitch = L.pitch_for(self.NY)
        r0, r1 = rows if rows is not None else (0, self.NX)
        self.geom = L.Geom(self.NX, self.NY, self.pitch, r0, r1, L.KT_F64, num_sms)
        self.pre, self.xr_unit, self.fused = bool(pre), bool(xr_unit), bool(fused)   # fused: BICGSolver.fuse_p_update
        out = (C.c_longlong * 8)()
        L.check(lib.kt_resident_plan(C.byref(self.geom), int(is_cg), int(pre), int(warm), int(enable), out))
        self.path, self.C, self.rows_per, self.NT, self.EPT = (int(v) for v in out[:5])
        if rows is not None:
            assert self.path == 0, 'slabs only run the streaming kernels'
        self._cache = {}
        self._lib, self._L = lib, L

    # ---- the summation itself -------------------------------------------------------------------------------------
    def __call__(self, kind, prod):
        prod = np.asarray(prod, dtype=np.float64)
        assert prod.shape == (self.NX, self.NY), (prod.shape, (self.NX, self.NY))
        if self.path == 0:
            return self._stream(kind, prod)
        if self.path in (1, 2):
            return self._cluster(prod)
        return self._coop(prod)

    # ---- streaming kernels ----------------------------------------------------------------------------------------
    def _plan(self, kind):
        if kind == 'end' and self.xr_unit:
            kind = 'end_unit'
        if kind == 'alpha' and self.fused:
            kind = 'alpha_fused'
        key = PLAN[kind]
        if key not in self._cache:
            out = (C.c_longlong * 12)()
            self._L.check(self._lib.kt_reduce_plan(C.byref(self.geom), key, int(self.pre), out))
            o = [int(v) for v in out]
            grid, block, V = o[1], o[2], o[3]
            assert block == 256
            if o[0] == 0:
                idx = self._ew_index(grid, V, o[4], o[5])
            else:
                idx = self._stencil_index(grid, V, W=o[4], TI=o[5], nCC=o[6], nTiles=o[7], row_lo=o[8], row_hi=o[9])
            self._cache[key] = (grid, _Gather(idx))
        return self._cache[key]

    def _ew_index(self, grid, V, v_lo, v_hi):
        # csrc/kt_krylov.cu KT_EW_OUTER / KT_EW_FOR: thread g walks the vectors v_lo + g, + stride, + 2 stride, ...
        stride = grid * 256
        n_steps = max(1, -(-(v_hi - v_lo) // stride))
        g = np.arange(stride, dtype=np.int64)[:, None, None]
        m = np.arange(n_steps, dtype=np.int64)[None, :, None]
        k = np.arange(V, dtype=np.int64)[None, None, :]
        vi = v_lo + g + m * stride
        idx = np.where(vi < v_hi, vi * V + k, -1)
        return idx.reshape(grid, 256, n_steps * V)

    def _stencil_index(self, grid, V, W, TI, nCC, nTiles, row_lo, row_hi):
        # csrc/kt_krylov.cu k_stencil: warp gw strides over the tiles; a tile is TI rows x (W x 32 vectors); per row the W
        # chunks in order, per chunk the V reals of this lane's vector
        nwarps = grid * 8
        per_warp = max(1, -(-nTiles // nwarps))
        gw = np.arange(nwarps, dtype=np.int64)[:, None, None, None, None, None]
        lane = np.arange(32, dtype=np.int64)[None, :, None, None, None, None]
        n = np.arange(per_warp, dtype=np.int64)[None, None, :, None, None, None]
        ri = np.arange(TI, dtype=np.int64)[None, None, None, :, None, None]
        w = np.arange(W, dtype=np.int64)[None, None, None, None, :, None]
        k = np.arange(V, dtype=np.int64)[None, None, None, None, None, :]
        tile = gw + n * nwarps
        rchunk, cc = tile // nCC, tile % nCC
        i = row_lo + rchunk * TI + ri
        j = ((cc * W + w) * 32 + lane) * V + k
        ok = (tile < nTiles) & (i < row_hi) & (j < self.pitch)
        idx = np.where(ok, i * self.pitch + j, -1)
        return idx.reshape(grid, 8, 32, per_warp * TI * W * V).reshape(grid, 256, -1)

    def _stream(self, kind, prod):
        grid, gather = self._plan(kind)
        plane = np.zeros((self.NX, self.pitch))
        plane[:, :self.NY] = prod
        t = gather.thread_sums(plane.ravel())                     # (grid, 256)
        partials = seq(tree(t.reshape(grid, 8, 32)))              # warp tree, then the 8 warps in order
        # finalising block: thread k adds partials k, k + 256, ... ; then the same two stages
        nb = -(-grid // 256) * 256
        f = seq(_pad(partials, nb).reshape(-1, 256).T.copy())     # (256,)
        return float(seq(tree(f.reshape(8, 32))[None, :])[0])

    # ---- cluster-resident solver ------------------------------------------------------------------------------------
    def _slices(self, prod, n_cta, rows_per, nt):
        """(n_cta, nt, L): thread t of CTA c owns the flat elements t, t + nt, ... of the CTA's compact (rows x NY) slice."""
        cap = rows_per * self.NY
        L = -(-cap // nt)
        rows = np.zeros((n_cta * rows_per, self.NY))
        rows[:self.NX] = prod                                   # CTA c holds rows [c rows_per, (c + 1) rows_per) (clipped to NX)
        out = np.zeros((n_cta, L * nt))
        out[:, :cap] = rows.reshape(n_cta, cap)
        return out.reshape(n_cta, L, nt).transpose(0, 2, 1)

    def _cluster(self, prod):
        Cn, rows_per, NT = self.C, self.rows_per, self.NT
        if self.path == 1:
            per_thread = seq(self._slices(prod, Cn, rows_per, NT))            # (C, NT)
        else:
            # k_cluster_solve_smem CL_FOR: thread (tx, ty) walks j = tx, tx + TC, ... (outer) and li = ty, ty + TR, ... (inner)
            NT = 512
            TC = min(self.NY, 512)
            TR = 512 // TC
            per_thread = np.zeros((Cn, NT))
            for c in range(Cn):
                r0, r1 = min(c * rows_per, self.NX), min((c + 1) * rows_per, self.NX)
                sl = prod[r0:r1]
                nrows = r1 - r0
                nj, nl = -(-self.NY // TC), -(-max(nrows, 1) // TR)
                buf = np.zeros((nl * TR, nj * TC))
                buf[:nrows, :self.NY] = sl
                # element (li = ty + a TR, j = tx + b TC) -> thread ty * TC + tx, position b * nl + a
                b4 = buf.reshape(nl, TR, nj, TC).transpose(1, 3, 2, 0).reshape(TR * TC, nj * nl)
                per_thread[c, :TR * TC] = seq(b4)
        warps = tree(per_thread.reshape(Cn, NT // 32, 32))                    # (C, NW)
        cta = tree(_pad(warps, 16))                                           # tree16 over the warp sums
        return float(tree(_pad(cta[None, :], 16))[0])                         # tree16 over the CTA slots

    # ---- cooperative solver -------------------------------------------------------------------------------------------
    def _coop(self, prod):
        G, rows_per = self.C, self.rows_per
        per_thread = seq(self._slices(prod, G, rows_per, 512))                # (G, 512)
        warps = tree(per_thread.reshape(G, 16, 32))                           # (G, 16)
        cta = tree(_pad(warps, 32))                                           # warp 0: 16 values + 16 zeros, butterfly
        lanes = seq(_pad(cta[None, :], 256).reshape(8, 32).T.copy())          # lane l: slots l, l + 32, ... in order
        return float(tree(lanes[None, :])[0])


class SlabOrder:
    """x-slab solve: every rank adds its owned rows in the streaming order, then the ranks are added in rank order
    (``k_p2p_reduce``; the NCCL path's all-reduce order is NCCL's own and is not restated).  ``pre``: the Jacobi variant."""

    def __init__(self, shape, bounds, is_cg=False, pre=False, num_sms=148):
        self.bounds = list(bounds)
        self.shape = shape
        self.ranks = []
        for r in range(len(self.bounds) - 1):
            g0 = self.bounds[r] - (1 if r > 0 else 0)
            g1 = self.bounds[r + 1] + (1 if r < len(self.bounds) - 2 else 0)
            l0 = self.bounds[r] - g0
            l1 = l0 + self.bounds[r + 1] - self.bounds[r]
            self.ranks.append((g0, g1, DeviceOrder((g1 - g0, shape[1]), is_cg=is_cg, pre=pre, enable=0, num_sms=num_sms, rows=(l0, l1),
                                                   xr_unit=True)))

    def __call__(self, kind, prod):
        d = 0.0
        for g0, g1, order in self.ranks:
            d = d + order(kind, prod[g0:g1])
        return float(d)


class SlabAutoOrder:
    """SIMPLE on x-slabs (``SlabKarmanSolver``): dispatches on the shape of the products to the ``SlabOrder`` of that system;
    ``bounds_by_rows`` maps the number of global rows of a plane to its slab boundaries.  The inlet / outlet column sums of
    ``correct_conserv`` live on one rank each and are added to zeros in rank order: the single-GPU order."""

    def __init__(self, bounds_by_rows, jacobi=True, num_sms=148):
        self.bounds_by_rows, self.jacobi, self.num_sms = dict(bounds_by_rows), jacobi, num_sms
        self._orders = {}

    def __call__(self, kind, prod):
        if kind == 'column':
            return column_sum(prod)
        key = (prod.shape, kind.startswith('cg_'))
        if key not in self._orders:
            self._orders[key] = SlabOrder(prod.shape, self.bounds_by_rows[prod.shape[0]], is_cg=key[1], pre=self.jacobi, num_sms=self.num_sms)
        return self._orders[key](kind, prod)


def column_sum(a):
    """``k_ch_outlet_scale`` (``csrc/kt_channel.cu``): one block of 256 threads, thread t adds the entries t, t + 256, ...
    of the inlet / outlet column (``karman_taichi.py:590-596``) in order -> lane tree -> the 8 warps in order."""
    a = np.asarray(a, dtype=np.float64).ravel()
    n = -(-max(a.size, 1) // 256) * 256
    t = seq(_pad(a[None, :], n).reshape(-1, 256).T.copy())
    return float(seq(tree(t.reshape(8, 32))[None, :])[0])


class AutoOrder:
    """Dispatches on the shape of the products and on the solver the inner product belongs to: what a SIMPLE run needs,
    where BiCGSTAB solves of the u and v systems and CG / BiCGSTAB solves of the p' system alternate.  ``jacobi`` = the
    monolith's solver (Jacobi plane + warm start: ``karman_taichi.py:342-456``)."""

    def __init__(self, jacobi=False, enable=1, num_sms=148):
        self.jacobi, self.enable, self.num_sms = bool(jacobi), enable, num_sms
        self._orders = {}

    def __call__(self, kind, prod):
        if kind == 'column':
            return column_sum(prod)
        is_cg = kind.startswith('cg_')
        key = (prod.shape, is_cg)
        if key not in self._orders:
            self._orders[key] = DeviceOrder(prod.shape, is_cg=is_cg, pre=self.jacobi, warm=self.jacobi, enable=self.enable,
                                            num_sms=self.num_sms)
        return self._orders[key](kind, prod)
