"""The peer-memory x-slab path with ALL ranks on ONE GPU: every rank is a thread with its own stream, the comm blocks and
halo-receive buffers are wired by address (`peers='local'`) instead of CUDA IPC.  The kernels are exactly the ones the
multi-GPU runs use -- boundary rows stored into the neighbour's receive buffer, flags, rank-ordered sums -- so this is what
exercises them on a one-GPU box (the driver's GPU test box has one; tests/test_gpu_slab.py::test_two_gpu_nccl needs two).

Every comparison is against the oracle adding in the slabs' order (oracle/device_order.py: SlabOrder / SlabAutoOrder): bit for bit.

Liveness caveat of this test set-up only: ranks that WAIT for each other while sharing one device rely on their streams making
progress side by side, which CUDA does not promise (one process per GPU -- the real deployment -- does not depend on it).  The
library removes the causes it can (kernels preloaded, no cudaFree meanwhile: slab.run_local_ranks); a run that still ends in
KT_STATUS_COMM_TIMEOUT (seen once per cold box, on its first process) is repeated from scratch -- new solvers, new flags -- up
to two more times.  Only the time-out is retried: a wrong number fails at once.
"""
import functools

import numpy as np
import pytest

from oracle import krylov
from oracle.device_order import SlabAutoOrder, SlabOrder, installed

pytestmark = pytest.mark.gpu


def retry_comm_timeout(test):
    @functools.wraps(test)
    def wrapper(*args, **kwargs):
        import gc
        import torch
        from karman_taichi_b200 import KtError
        for attempt in range(3):
            try:
                return test(*args, **kwargs)
            except KtError as e:
                if attempt == 2 or not ('COMM_TIMEOUT' in str(e) or 'gave up waiting' in str(e)):
                    raise
                print(f'{test.__name__}: attempt {attempt + 1} timed out waiting for a co-scheduled rank ({e}); repeating')
                torch.cuda.synchronize()
                gc.collect()
    return wrapper


def sms(cuda):
    import torch
    return torch.cuda.get_device_properties(cuda).multi_processor_count


def run_ranks(cuda, n, fn):
    """fn(rank) on n threads, each with its own stream (slab.run_local_ranks: nothing may free device memory meanwhile)."""
    from karman_taichi_b200.slab import run_local_ranks
    return run_local_ranks(n, fn, device=cuda, timeout=300)


def slab_solvers(cuda, coef, rhs, P):
    import karman_taichi_b200 as kt
    from karman_taichi_b200.slab import P2PSlabSolver, SlabPartition
    part = SlabPartition(rhs.shape[0], P)
    solvers = []
    for r in range(P):
        c_l = np.ascontiguousarray(part.scatter(coef, r))
        b_l = np.ascontiguousarray(part.scatter(rhs, r))
        mk = lambda a: kt.Field(a.shape, np.float64, cuda).from_numpy(a)
        solvers.append(P2PSlabSolver(mk(c_l), mk(b_l), kt.Field(b_l.shape, np.float64, cuda), part, r, peers='local'))
    P2PSlabSolver.connect_local(solvers)
    return part, solvers


@pytest.mark.parametrize('P,case', [(2, ('poisson', 96, 0.1)), (3, ('random', 130, 67)), (5, ('random', 37, 300))])
@retry_comm_timeout
def test_p2p_bicgstab_and_cg_threads_bit_exact(cuda, P, case):
    kind, a, b_ = case
    coef, rhs = krylov.poisson_sin_system(a, b_) if kind == 'poisson' else krylov.random_dd_system(a, b_, seed=0)
    part, solvers = slab_solvers(cuda, coef, rhs, P)
    for _ in range(2):                                            # twice: epochs / sequence numbers carry over between solves
        res = run_ranks(cuda, P, lambda r: solvers[r].solve(eps=1e-10, max_iter=rhs.size))
    assert len(set(res)) == 1                                     # identical (iterations, status) on every rank
    x = np.concatenate([part.owned_of_local(s.x.to_numpy(), r) for r, s in enumerate(solvers)], axis=0)
    with installed(SlabOrder(rhs.shape, part.bounds, num_sms=sms(cuda))):
        x_dev, it_dev, st_dev = krylov.bicgstab_solve(coef, rhs, eps=1e-10)
    assert res[0] == (it_dev, st_dev)
    np.testing.assert_array_equal(x, x_dev)
    if kind == 'poisson':
        res = run_ranks(cuda, P, lambda r: solvers[r].solve_cg(eps=1e-10, max_iter=rhs.size))
        xc = np.concatenate([part.owned_of_local(s.x.to_numpy(), r) for r, s in enumerate(solvers)], axis=0)
        with installed(SlabOrder(rhs.shape, part.bounds, is_cg=True, num_sms=sms(cuda))):
            xc_dev, itc_dev, stc_dev = krylov.cg_solve(coef, rhs, eps=1e-10)
        assert res[0] == (itc_dev, stc_dev) and len(set(res)) == 1
        np.testing.assert_array_equal(xc, xc_dev)


@retry_comm_timeout
def test_p2p_exchange_and_allreduce_threads(cuda):
    """kt_p2p_exchange / kt_p2p_allreduce on their own: many back-to-back calls (slot / group rotation), sums in rank order."""
    import torch
    coef, rhs = krylov.random_dd_system(61, 40, seed=2)
    P = 4
    part, solvers = slab_solvers(cuda, coef, rhs, P)
    planes = [torch.zeros((part.local_nx(r), solvers[r].b.pitch), dtype=torch.float64, device=cuda) for r in range(P)]
    vals = [0.1 * (r + 1) + 1e-17 * r for r in range(P)]

    def work(r):
        s = solvers[r]
        l0, l1 = part.local_owned(r)
        sums = []
        for k in range(12):
            planes[r][l0:l1] = float(100 * k + r)                 # my owned rows carry (round, rank)
            s.exchange(planes[r].data_ptr())
            lo = float(planes[r][l0 - 1, 3]) if r > 0 else None
            hi = float(planes[r][l1, 3]) if r < P - 1 else None
            assert lo in (None, 100 * k + r - 1) and hi in (None, 100 * k + r + 1), (r, k, lo, hi)
            sums.append(s.allreduce(vals[r] * (k + 1), float(r), 'sum'))
            assert s.allreduce(float(r * k), -1.0 - r, 'max') == (float((P - 1) * k), 0.0)   # d starts at 0.0 (max |p'| >= 0)
        return sums
    out = run_ranks(cuda, P, work)
    assert all(o == out[0] for o in out)                          # the same bits on every rank
    for k in range(12):
        d = 0.0
        for r in range(P):
            d = d + vals[r] * (k + 1)
        assert out[0][k] == (d, float(sum(range(P))))


@pytest.mark.parametrize('P,kw,nsub', [(2, dict(nx=48, ny=16, cyl=(12, 8, 10)), 3), (3, dict(nx=96, ny=48, cyl=(64, 32, 144)), 2),
                                       (4, dict(nx=320, ny=64, cyl=(64, 32, 144)), 1)])
@retry_comm_timeout
def test_slab_karman_threads_bit_exact(cuda, P, kw, nsub):
    """SlabKarmanSolver on P ranks (threads) against the oracle in the slabs' summation order: every field, every Krylov
    count, bit for bit; against the one-GPU KarmanSolver: assembly identical, fields to 1e-8."""
    import karman_taichi_b200 as kt
    from oracle.karman import KarmanOracle
    ranks = [kt.SlabKarmanSolver(r, P, device=cuda, peers='local', **kw) for r in range(P)]
    kt.SlabKarmanSolver.connect_local(ranks)
    o = KarmanOracle(**kw)
    one = kt.KarmanSolver(device=cuda, **kw)
    nx, ny = kw['nx'], kw['ny']
    order = SlabAutoOrder({nx + 3: ranks[0].part_u.bounds, nx + 2: ranks[0].part_p.bounds}, num_sms=sms(cuda))

    def gather(name):
        return np.concatenate([s.owned(getattr(s, name)) for s in ranks], axis=0)

    run_ranks(cuda, P, lambda r: ranks[r].init())
    o.init(); one.init()
    for name in ('u', 'v', 'p'):
        np.testing.assert_array_equal(gather(name), getattr(o, name), err_msg=name)
    ct = np.concatenate([s.ct[s.part_p.local_owned(s.rank)[0]:s.part_p.local_owned(s.rank)[1], :ny + 2].cpu().numpy() for s in ranks], axis=0)
    np.testing.assert_array_equal(ct, o.ct)
    # first assembly: the same bits as on one GPU (no sums involved)
    run_ranks(cuda, P, lambda r: ranks[r].fill_Au())
    one.fill_Au()
    for name in ('coef_u', 'b_u', 'M_u', 'Minv_u', 'coef_v', 'b_v', 'M_v'):
        np.testing.assert_array_equal(gather(name), getattr(one, name).to_numpy(), err_msg=name)
    with installed(order):
        for k in range(nsub):
            po = o.sub_iteration(1e-8)
            pg = run_ranks(cuda, P, lambda r: (ranks[r].sub_iteration(1e-8), dict(ranks[r].last_iters), dict(ranks[r].last_status)))
            assert all(x == pg[0] for x in pg)                    # same max |p'|, counts and statuses on every rank
            assert pg[0] == (po, o.last_iters, o.last_status), k
    for name in ('u', 'v', 'p', 'pcor', 'udiv', 'coef_p', 'b_p'):
        np.testing.assert_array_equal(gather(name), getattr(o, name), err_msg=name)
    # against the one-GPU solver: different partition of the sums only
    for k in range(nsub):
        one.sub_iteration(1e-8)
    for name in ('u', 'p'):
        a, b = gather(name), getattr(one, name).to_numpy()
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-8, name
