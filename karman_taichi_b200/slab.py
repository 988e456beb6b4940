"""x-slab domain decomposition of the BiCGSTAB solve across the GPUs of one box (north_star (e); no reference
counterpart -- the reference is a single address space).

One process per GPU (``torch.distributed``: NCCL on GPUs, gloo in the CPU tests).  The ``(NX, NY)`` planes are split
along ``i`` (the slowest axis) into contiguous row blocks; a rank stores its owned rows plus one halo row on each
interior side, so a halo is one contiguous (pitched) row.  Per BiCGSTAB iteration the ranks exchange the boundary
rows of ``p`` (before ``v = A p``), of ``v`` and of ``r`` (before the fused ``s = r - alpha v, t = A s``), and
all-reduce the 1-2 local inner products of each of the three fused kernels; every rank then runs the same scalar
algebra on the identical reduced values, so all ranks take the same convergence decision without talking again.

``DistBiCGSTAB`` is the host-side driver; the arithmetic lives behind an ``ops`` object:
``CudaSlabOps`` (libkarman_b200 ``kt_dist_*`` phases -- the product) or, in ``tests/``, a NumPy stand-in built on
the oracle, which is how the partition / halo / reduction logic is tested on CPU with gloo.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

from . import _lib

FIN_NONE, FIN_BICG_INIT, FIN_BICG_ALPHA, FIN_BICG_OMEGA, FIN_BICG_END = 0, 1, 2, 3, 4
FIN_CG_INIT, FIN_CG_ALPHA, FIN_CG_END = 5, 6, 7


class SlabPartition:
    """Rows ``[0, NX)`` of the global plane split into ``nranks`` contiguous blocks (sizes differ by <= 1)."""

    def __init__(self, NX, nranks):
        if nranks < 1 or NX < nranks:
            raise ValueError('need 1 <= nranks <= NX')
        self.NX, self.nranks = int(NX), int(nranks)
        base, rem = divmod(self.NX, self.nranks)
        self.bounds = [0]
        for r in range(self.nranks):
            self.bounds.append(self.bounds[-1] + base + (1 if r < rem else 0))

    @classmethod
    def from_bounds(cls, bounds):
        """Explicit block boundaries ``[0, b1, ..., NX]`` (the u-type planes of the staggered grid have one row more than the
        cell-centred ones and must still be cut at the same rows)."""
        bounds = [int(b) for b in bounds]
        if bounds[0] != 0 or any(b1 <= b0 for b0, b1 in zip(bounds, bounds[1:])):
            raise ValueError('bounds must start at 0 and increase strictly')
        self = cls.__new__(cls)
        self.NX, self.nranks, self.bounds = bounds[-1], len(bounds) - 1, bounds
        return self

    def owned(self, rank):
        """Global row range owned by ``rank``."""
        return self.bounds[rank], self.bounds[rank + 1]

    def halo(self, rank):
        """(rows of halo below, rows of halo above) = (1 if there is a lower neighbour, 1 if an upper one)."""
        return (1 if rank > 0 else 0), (1 if rank < self.nranks - 1 else 0)

    def stored(self, rank):
        """Global row range present in the rank's local plane (owned rows + halos)."""
        lo, hi = self.owned(rank)
        hl, hu = self.halo(rank)
        return lo - hl, hi + hu

    def local_owned(self, rank):
        """Owned rows in local indices: what ``kt_set_rows`` gets."""
        lo, hi = self.owned(rank)
        hl, _ = self.halo(rank)
        return hl, hl + (hi - lo)

    def local_nx(self, rank):
        a, b = self.stored(rank)
        return b - a

    def scatter(self, global_array, rank):
        """Local copy (owned rows + halo rows filled from the global array) of a NumPy / torch array whose first
        axis is ``i``."""
        a, b = self.stored(rank)
        return global_array[a:b]

    def owned_of_local(self, local_array, rank):
        l0, l1 = self.local_owned(rank)
        return local_array[l0:l1]


class HaloExchanger:
    """Boundary-row exchange with the two x-neighbours over ``torch.distributed`` point-to-point ops."""

    def __init__(self, part: SlabPartition, rank, group=None):
        self.part, self.rank, self.group = part, rank, group
        self.l0, self.l1 = part.local_owned(rank)
        self.has_lo, self.has_hi = (h == 1 for h in part.halo(rank))

    def exchange(self, plane: torch.Tensor):
        """``plane``: the local ``(NX_local, pitch)`` tensor.  Sends the first / last owned row to the lower /
        upper neighbour and receives their boundary rows into this rank's halo rows."""
        ops = []
        if self.has_lo:
            ops.append(dist.P2POp(dist.isend, plane[self.l0], self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, plane[self.l0 - 1], self.rank - 1, self.group))
        if self.has_hi:
            ops.append(dist.P2POp(dist.isend, plane[self.l1 - 1], self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, plane[self.l1], self.rank + 1, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class CudaSlabOps:
    """The product arithmetic of one rank: ``kt_dist_*`` phases of libkarman_b200 on this rank's slab."""

    def __init__(self, coef, b, x, part: SlabPartition, rank):
        from .cgsolver import _Handle
        from .field import Field
        self.coef, self.b, self.x, self.part = coef, b, x, part
        self.lib = _lib.load()
        self._hd = _Handle(b)
        l0, l1 = part.local_owned(rank)
        _lib.check(self.lib.kt_set_rows(self._hd.h, l0, l1))
        mk = lambda: Field(b.shape, b.np_dtype, b.device)
        self.r, self.p, self.v, self.s, self.t = mk(), mk(), mk(), mk(), mk()
        self.dots = torch.zeros(2, dtype=torch.float64, device=b.device)

    @property
    def h(self):
        self._hd.use_current_stream()
        return self._hd.h

    def _phase(self, phase, a0=None, a1=None, a2=None, o0=None, o1=None, o2=None):
        P = lambda f: C.c_void_p(f.ptr) if f is not None else None
        _lib.check(self.lib.kt_dist_bicg_phase(self.h, phase, C.c_void_p(self.coef.ptr), P(a0), P(a1), P(a2), P(o0), P(o1), P(o2),
                                               C.c_void_p(self.dots.data_ptr())))

    def planes(self):
        return {'p': self.p.data, 'v': self.v.data, 'r': self.r.data}

    def default_max_iter(self):
        return self.part.NX * self.b.NY

    def begin(self, eps, max_iter):
        _lib.check(self.lib.kt_dist_begin(self.h, float(eps), int(max_iter), 0))

    def init(self):
        self._phase(0, a0=self.b, o0=self.r, o1=self.x)          # r_tld aliases b (x0 = 0)

    def update_p(self):
        self._phase(1, a0=self.r, a1=self.v, o0=self.p)

    def v_dot(self):
        self._phase(2, a0=self.p, a1=self.b, o0=self.v)

    def st_dots(self):
        self._phase(3, a0=self.r, a1=self.v, o0=self.s, o1=self.t)

    def xr_dots(self):
        self._phase(4, a0=self.p, a1=self.s, a2=self.t, o0=self.x, o1=self.r, o2=self.b)

    # CG phases (cgsolver.py:83-114); Ap lives in the v plane
    def cg_init(self):
        self._phase(5, a0=self.b, o0=self.r, o1=self.x, o2=self.p)

    def cg_ap_dot(self):
        self._phase(6, a0=self.p, o0=self.v)

    def cg_xr_dot(self):
        self._phase(7, a0=self.p, a1=self.v, o0=self.x, o1=self.r)

    def cg_update_p(self):
        self._phase(8, a0=self.r, o0=self.p)

    def scalar(self, fin):
        _lib.check(self.lib.kt_dist_scalar(self.h, fin, C.c_void_p(self.dots.data_ptr()), None, 0))

    def poll(self):
        done, it, st, rel = C.c_int(), C.c_long(), C.c_int(), C.c_double()
        _lib.check(self.lib.kt_dist_poll(self.h, C.byref(done), C.byref(it), C.byref(st), C.byref(rel)))
        return bool(done.value), it.value, st.value, rel.value

    def launch_count(self):
        return self._hd.launch_count()


class DistBiCGSTAB:
    """``BICGSolver.solve`` (``bicgsolver.py:88-135``) over x-slabs: same recurrence, same stopping test."""

    def __init__(self, ops, exchanger: HaloExchanger, group=None, poll_every=8):
        self.ops, self.ex, self.group, self.poll_every = ops, exchanger, group, poll_every
        self.iterations, self.status, self.rel_res = 0, 0, float('nan')

    def _allreduce(self):
        dist.all_reduce(self.ops.dots, op=dist.ReduceOp.SUM, group=self.group)

    def solve(self, eps=1e-8, max_iter=None, fixed_iters=False):
        ops, ex = self.ops, self.ex
        planes = ops.planes()
        if max_iter is None:
            max_iter = ops.default_max_iter()          # the reference's `steps` = NX * NY of the whole grid (cgsolver.py:14)
        ops.begin(-1.0 if fixed_iters else eps, max_iter)
        ops.init()
        self._allreduce()
        ops.scalar(FIN_BICG_INIT)
        ex.exchange(planes['r'])
        it = 0
        while it < max_iter:
            ops.update_p()
            ex.exchange(planes['p'])
            ops.v_dot()
            self._allreduce()
            ops.scalar(FIN_BICG_ALPHA)
            ex.exchange(planes['v'])
            ops.st_dots()
            self._allreduce()
            ops.scalar(FIN_BICG_OMEGA)
            ops.xr_dots()
            self._allreduce()
            ops.scalar(FIN_BICG_END)
            ex.exchange(planes['r'])
            it += 1
            if not fixed_iters and it % self.poll_every == 0:
                done, _, _, _ = ops.poll()
                if done:
                    break
        done, self.iterations, self.status, self.rel_res = ops.poll()
        return self.iterations, self.status


class DistCG:
    """``CGSolver.solve`` (``cgsolver.py:83-114``) over x-slabs: one halo exchange (``p``) and two all-reduces per
    iteration; same recurrence and stopping test (``sqrt(rTr) < sqrt(rTr0) * eps``, ``:102``)."""

    def __init__(self, ops, exchanger: HaloExchanger, group=None, poll_every=8):
        self.ops, self.ex, self.group, self.poll_every = ops, exchanger, group, poll_every
        self.iterations, self.status, self.rel_res = 0, 0, float('nan')

    def _allreduce(self):
        dist.all_reduce(self.ops.dots, op=dist.ReduceOp.SUM, group=self.group)

    def solve(self, eps=1e-8, max_iter=None, fixed_iters=False):
        ops, ex = self.ops, self.ex
        p = ops.planes()['p']
        if max_iter is None:
            max_iter = ops.default_max_iter()
        ops.begin(-1.0 if fixed_iters else eps, max_iter)
        ops.cg_init()                       # r = b, x = 0, p = r (beta = 0: cgsolver.py:90)
        self._allreduce()
        ops.scalar(FIN_CG_INIT)
        ex.exchange(p)
        it = 0
        while it < max_iter:
            ops.cg_ap_dot()
            self._allreduce()
            ops.scalar(FIN_CG_ALPHA)
            ops.cg_xr_dot()
            self._allreduce()
            ops.scalar(FIN_CG_END)
            ops.cg_update_p()
            ex.exchange(p)
            it += 1
            if not fixed_iters and it % self.poll_every == 0:
                done, _, _, _ = ops.poll()
                if done:
                    break
        done, self.iterations, self.status, self.rel_res = ops.poll()
        return self.iterations, self.status


# ---------------------------------------------------------------------------------------------------------------
# Peer-memory (NVLink) variant: no NCCL in the solver loop
# ---------------------------------------------------------------------------------------------------------------
class _IpcBuffer:
    """A zero-initialised device buffer owned by the library (dedicated cudaMalloc) that peer ranks can map."""

    def __init__(self, hd, nbytes):
        self.hd, self.nbytes = hd, int(nbytes)
        ptr = C.c_void_p()
        self.handle = C.create_string_buffer(64)
        _lib.check(hd.lib.kt_ipc_alloc(hd.h, self.nbytes, C.byref(ptr), self.handle))
        self.ptr = ptr.value

    def handle_bytes(self):
        return bytes(self.handle.raw)

    def __del__(self):
        try:
            if self.ptr and self.hd.h:
                self.hd.lib.kt_ipc_free(self.hd.h, C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def run_local_ranks(n, fn, device=None, timeout=600.0):
    """``fn(rank)`` for ``rank = 0 .. n-1`` on ``n`` host threads with one CUDA stream each: how an in-process multi-rank run
    (solvers built with ``peers='local'``, all ranks on ONE device) is driven.  Returns the list of results.

    The comm kernels of different ranks wait for each other, so while the threads run nothing may force the device idle:

    * every kernel is loaded beforehand (``kt_preload_kernels``, called by ``P2PSlabSolver``): CUDA's lazy loader waits for
      running kernels;
    * no device memory may be freed: ``cudaFree`` -- a handle or an IPC buffer of an earlier solver reaching its ``__del__``
      through the cyclic garbage collector in the middle of a solve -- waits for ALL kernels on the device, including the one
      that spins for the thread that called it.  Garbage is therefore collected first and the collector is off until every
      rank has returned.

    One process per GPU (``peers='ipc'``) needs none of this: every process owns its device."""
    import gc
    from concurrent.futures import ThreadPoolExecutor
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
    streams = [torch.cuda.Stream(device) for _ in range(n)]
    torch.cuda.synchronize(device)

    def wrapped(r):
        with torch.cuda.device(device), torch.cuda.stream(streams[r]):
            out = fn(r)
            streams[r].synchronize()
            return out

    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    watchdog = float(os.environ.get('KT_LOCAL_RANKS_WATCHDOG', '0') or 0)   # seconds: dump every thread's stack if still running
    if watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(watchdog, exit=False, file=sys.stderr)
    try:
        with ThreadPoolExecutor(max_workers=n) as ex:
            futs = [ex.submit(wrapped, r) for r in range(n)]
            return [f.result(timeout=timeout) for f in futs]
    finally:
        if watchdog > 0:
            faulthandler.cancel_dump_traceback_later()
        if was_enabled:
            gc.enable()


class P2PSlabSolver:
    """x-slab BiCGSTAB / CG whose halo exchange and inner-product all-reduce run inside the library over NVLink peer memory
    (``kt_p2p_bicgstab_solve`` / ``kt_p2p_cg_solve``).  Every rank maps its two neighbours' small halo-receive buffers
    (3 planes x 2 sides x one pitched row) and every rank's comm block through CUDA IPC once, at construction (the only
    collective use of ``torch.distributed`` here: exchanging the 64-byte IPC handles); the field planes are ordinary
    allocations nobody else maps.  After that a solve is one C-ABI call per rank and no host code runs per iteration.

    ``Minv`` / ``warm_start`` / ``omega_zero_break``: the monolith's Jacobi-preconditioned, warm-started variant
    (``karman_taichi.py:342-456``), as on ``BICGSolver``; the ``Minv`` plane must be valid on the halo rows as well (the operand of
    the product is ``Minv * p``: exchange it after assembling it, ``SlabKarmanSolver.fill_Au``).  ``ipc_planes=True`` (measurement aid) puts r / p / v themselves
    into IPC-exported allocations, which is how round 1 ran (profiles/r02_p2p.md)."""

    def __init__(self, coef, b, x, part: SlabPartition, rank, group=None, ipc_planes=False, peers='ipc'):
        from .cgsolver import _Handle
        from .field import Field
        self.coef, self.b, self.x, self.part, self.rank = coef, b, x, part, rank
        self.lib = _lib.load()
        self._hd = _Handle(b)
        l0, l1 = part.local_owned(rank)
        _lib.check(self.lib.kt_set_rows(self._hd.h, l0, l1))
        # the comm kernels wait for kernels of other streams: nothing may be left to the lazy module loader (kt_preload_kernels)
        if int(self.lib.kt_preload_kernels(self._hd.h)) <= 0:
            _lib.check(-4)
        self.t = Field(b.shape, b.np_dtype, b.device)       # s is recomputed by the x/r kernel, never stored
        esize = b.data.element_size()
        if ipc_planes:
            plane_bytes = b.data.numel() * esize
            self._planes = [_IpcBuffer(self._hd, plane_bytes) for _ in range(3)]
            self._r, self._p, self._v = (pl.ptr for pl in self._planes)
        else:
            self._planes = [Field(b.shape, b.np_dtype, b.device) for _ in range(3)]
            self._r, self._p, self._v = (pl.ptr for pl in self._planes)
        self._r_tld = None
        self.Minv, self.warm_start, self.omega_zero_break = None, False, False
        # what neighbours write into: boundary rows (3 planes x 2 sides x pitch) and every rank's sums / flags
        self._halo = _IpcBuffer(self._hd, int(self.lib.kt_p2p_halo_slots()) * 2 * b.pitch * esize)
        self._blk = _IpcBuffer(self._hd, int(self.lib.kt_p2p_comm_block_bytes()))
        self.dots = torch.zeros(2, dtype=torch.float64, device=b.device)
        self.red_in = torch.zeros(2, dtype=torch.float64, device=b.device)
        self.iterations, self.status, self.rel_res = 0, 0, float('nan')
        self._mapped = []
        n = part.nranks
        if peers == 'local':
            # all ranks live in THIS process (one thread and one stream each, any number of them on one GPU): the buffers are
            # reachable by their own addresses -- connect_local() wires them once every rank's solver exists.  This is how the
            # real comm kernels are tested on a single-GPU box.
            self.comm = None
            return
        mine = {'halo': self._halo.handle_bytes(), 'blk': self._blk.handle_bytes()}
        descs = [None] * n
        if n > 1:
            dist.all_gather_object(descs, mine, group=group)
        else:
            descs[0] = mine

        def open_(hb):
            out = C.c_void_p()
            _lib.check(self.lib.kt_ipc_open(self._hd.h, hb, C.byref(out)))
            self._mapped.append(out.value)
            return out.value

        err = None
        try:
            cm = _lib.P2PComm()
            cm.rank, cm.nranks = rank, n
            for k in range(16):
                cm.comm_block[k] = None
            for r_ in range(n):
                cm.comm_block[r_] = self._blk.ptr if r_ == rank else open_(descs[r_]['blk'])
            for side, nb in ((0, rank - 1), (1, rank + 1)):
                cm.peer_halo_recv[side] = open_(descs[nb]['halo']) if 0 <= nb < n else None
            cm.halo_recv = self._halo.ptr
            cm.local_dots = self.dots.data_ptr()
            self.comm = cm
        except Exception as e:      # e.g. no peer access between two devices: every rank must learn of it (no hang)
            err = e
        if n > 1:
            # every rank has mapped everything before anyone starts signalling; a rank that could not map tells the others
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=b.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                raise _lib.KtError(f'x-slab peer-memory setup failed on rank {rank}: {err}' if err else
                                   'x-slab peer-memory setup failed on another rank')
        elif err:
            raise err

    @staticmethod
    def connect_local(solvers):
        """Wire the comm structs of P solvers built with ``peers='local'`` (index = rank)."""
        n = len(solvers)
        for r, s in enumerate(solvers):
            cm = _lib.P2PComm()
            cm.rank, cm.nranks = r, n
            for k in range(16):
                cm.comm_block[k] = solvers[k]._blk.ptr if k < n else None
            cm.peer_halo_recv[0] = solvers[r - 1]._halo.ptr if r > 0 else None
            cm.peer_halo_recv[1] = solvers[r + 1]._halo.ptr if r < n - 1 else None
            cm.halo_recv = s._halo.ptr
            cm.local_dots = s.dots.data_ptr()
            s.comm = cm

    def exchange(self, plane_ptr):
        """Boundary rows of one plane of this system's geometry to / from the two neighbours (``kt_p2p_exchange``)."""
        self._hd.use_current_stream()
        _lib.check(self.lib.kt_p2p_exchange(self._hd.h, C.byref(self.comm), C.c_void_p(plane_ptr)))

    def allreduce(self, a, b=0.0, op='sum', device_pair=None):
        """Two doubles reduced over the ranks in rank order (``kt_p2p_allreduce``); returns the pair."""
        self._hd.use_current_stream()
        out = (C.c_double * 2)()
        if device_pair is None:       # host values travel as kernel arguments (no host-to-device copy in the stream)
            vals = (C.c_double * 2)(float(a), float(b))
            _lib.check(self.lib.kt_p2p_allreduce(self._hd.h, C.byref(self.comm), None, vals, 1 if op == 'max' else 0, out))
        else:
            _lib.check(self.lib.kt_p2p_allreduce(self._hd.h, C.byref(self.comm), C.c_void_p(device_pair.data_ptr()), None,
                                                 1 if op == 'max' else 0, out))
        return out[0], out[1]

    def __del__(self):
        try:
            for m in getattr(self, '_mapped', []):
                self.lib.kt_ipc_close(self._hd.h, C.c_void_p(m))
            self._mapped = []
        except Exception:
            pass

    def launch_count(self):
        return self._hd.launch_count()

    def _ws(self, t=True):
        r_tld = None
        if self.warm_start:
            if self._r_tld is None:
                from .field import Field
                self._r_tld = Field(self.b.shape, self.b.np_dtype, self.b.device)
            r_tld = self._r_tld.ptr
        return _lib.KrylovWs(r=self._r, p=self._p, Ap=self._v, r_tld=r_tld, s=None, t=(self.t.ptr if t else None), history=None,
                             history_cap=0, p2=None, Ap2=None)

    def solve(self, eps=1e-8, max_iter=None, fixed_iters=False):
        self._hd.use_current_stream()
        ws = self._ws()
        it, rel, st = C.c_long(), C.c_double(), C.c_int()
        if max_iter is None:
            max_iter = self.part.NX * self.b.NY
        flags = (_lib.FLAG_WARM_START if self.warm_start else 0) | (_lib.FLAG_OMEGA_ZERO_BREAK if self.omega_zero_break else 0)
        minv = C.c_void_p(self.Minv.ptr) if self.Minv is not None else None
        _lib.check(self.lib.kt_p2p_bicgstab_solve(self._hd.h, C.c_void_p(self.coef.ptr), C.c_void_p(self.b.ptr), C.c_void_p(self.x.ptr), minv,
                                                  C.byref(ws), C.byref(self.comm), float(eps), int(max_iter), flags, 1 if fixed_iters else 0,
                                                  C.byref(it), C.byref(rel), C.byref(st)))
        return self._finish(it, rel, st)

    def _finish(self, it, rel, st):
        self.iterations, self.status, self.rel_res = it.value, st.value, rel.value
        if st.value == _lib.STATUS_COMM_TIMEOUT:
            raise _lib.KtError('x-slab peer-memory solve: a neighbour rank never signalled (KT_STATUS_COMM_TIMEOUT): '
                               + (self.lib.kt_last_error() or b'').decode(errors='replace'))
        return self.iterations, self.status

    def solve_cg(self, eps=1e-8, max_iter=None, fixed_iters=False):
        """``CGSolver.solve`` over the same slabs and peer mappings (``kt_p2p_cg_solve``): Ap lives in the v plane."""
        self._hd.use_current_stream()
        ws = self._ws(t=False)
        it, rel, st = C.c_long(), C.c_double(), C.c_int()
        if max_iter is None:
            max_iter = self.part.NX * self.b.NY
        _lib.check(self.lib.kt_p2p_cg_solve(self._hd.h, C.c_void_p(self.coef.ptr), C.c_void_p(self.b.ptr), C.c_void_p(self.x.ptr),
                                            C.byref(ws), C.byref(self.comm), float(eps), int(max_iter), 1 if fixed_iters else 0,
                                            C.byref(it), C.byref(rel), C.byref(st)))
        return self._finish(it, rel, st)
