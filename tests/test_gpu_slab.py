"""GPU parity of the x-slab (multi-GPU) phases.

* ``test_emulated_slabs_single_gpu``: P virtual ranks on ONE device -- every rank's slab goes through the real
  ``kt_dist_*`` kernels (row-range restricted SpMV / updates / dots), the halo exchange and the all-reduce are done by
  hand in the test.  P-slab result must equal the 1-slab result.
* ``test_two_gpu_nccl``: the real thing over NCCL when the box has >= 2 GPUs (skipped otherwise).
"""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import krylov

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize('P,case', [(2, ('poisson', 96, 0.1, 1e-10)), (3, ('random', 130, 67, 1e-10)), (4, ('random', 37, 300, 1e-9))])
def test_emulated_slabs_single_gpu(cuda, P, case):
    import torch
    import karman_taichi_b200 as kt
    from karman_taichi_b200.slab import CudaSlabOps, SlabPartition
    kind, a, b_, eps = case
    coef, rhs = krylov.poisson_sin_system(a, b_) if kind == 'poisson' else krylov.random_dd_system(a, b_, seed=0)
    x_ref, it_ref, st_ref = krylov.bicgstab_solve(coef, rhs, eps=eps)
    part = SlabPartition(rhs.shape[0], P)
    ranks = []
    for r in range(P):
        c_l = np.ascontiguousarray(part.scatter(coef, r))
        b_l = np.ascontiguousarray(part.scatter(rhs, r))
        fc = kt.Field(c_l.shape, np.float64, cuda).from_numpy(c_l)
        fb = kt.Field(b_l.shape, np.float64, cuda).from_numpy(b_l)
        fx = kt.Field(b_l.shape, np.float64, cuda)
        ranks.append(CudaSlabOps(fc, fb, fx, part, r))

    def exchange(name):
        for r in range(P - 1):
            lo, up = ranks[r].planes()[name], ranks[r + 1].planes()[name]
            l0r, l1r = part.local_owned(r)
            l0u, _ = part.local_owned(r + 1)
            lo[l1r].copy_(up[l0u])           # upper neighbour's first owned row -> my top halo
            up[l0u - 1].copy_(lo[l1r - 1])   # my last owned row -> its bottom halo

    def allreduce():
        tot = sum(o.dots for o in ranks)
        for o in ranks:
            o.dots.copy_(tot)

    def each(fn):
        for o in ranks:
            getattr(o, fn)()

    def scalar(fin):
        for o in ranks:
            o.scalar(fin)

    for o in ranks:
        o.begin(eps, rhs.size)
    each('init'); allreduce(); scalar(1); exchange('r')
    for it in range(rhs.size):
        each('update_p'); exchange('p')
        each('v_dot'); allreduce(); scalar(2); exchange('v')
        each('st_dots'); allreduce(); scalar(3)
        each('xr_dots'); allreduce(); scalar(4); exchange('r')
        polls = [o.poll() for o in ranks]
        assert len({(p[0], p[1], p[2]) for p in polls}) == 1      # identical decisions on every rank
        if polls[0][0]:
            break
    done, iters, status, rel = polls[0]
    assert status == st_ref == 0
    x = np.concatenate([part.owned_of_local(o.x.to_numpy(), r) for r, o in enumerate(ranks)], axis=0)
    # order pinned (each slab's owned rows in the streaming kernels' order, then the ranks in rank order): bit for bit
    from oracle.device_order import SlabOrder, installed
    with installed(SlabOrder(rhs.shape, part.bounds, num_sms=torch.cuda.get_device_properties(cuda).multi_processor_count)):
        x_dev, it_dev, st_dev = krylov.bicgstab_solve(coef, rhs, eps=eps)
    assert (iters, status) == (it_dev, st_dev)
    np.testing.assert_array_equal(x, x_dev)
    assert abs(iters - it_ref) <= max(3, round(0.1 * it_ref))      # NumPy's pairwise order: tests/test_gpu_krylov.py
    assert rel_l2(x, x_ref) < 1e-8
    # and against the single-slab CUDA solve: same kernels, different partition of the sums
    fc = kt.Field(coef.shape, np.float64, cuda).from_numpy(coef)
    fb = kt.Field(rhs.shape, np.float64, cuda).from_numpy(rhs)
    fx = kt.Field(rhs.shape, np.float64, cuda)
    s = kt.BICGSolver(fc, fb, fx)
    s.solve(eps=eps)
    assert rel_l2(x, fx.to_numpy()) < 1e-8


@pytest.mark.parametrize('P,n,offset', [(2, 96, 0.1), (3, 61, 0.0), (5, 130, 0.1)])
def test_emulated_slabs_cg_single_gpu(cuda, P, n, offset):
    """CGSolver.solve (cgsolver.py:83-114) over P virtual ranks on one device: phases 5-8 of kt_dist_bicg_phase."""
    import karman_taichi_b200 as kt
    from karman_taichi_b200.slab import CudaSlabOps, SlabPartition
    coef, rhs = krylov.poisson_sin_system(n, offset)
    eps = 1e-10
    x_ref, it_ref, st_ref = krylov.cg_solve(coef, rhs, eps=eps)
    part = SlabPartition(rhs.shape[0], P)
    ranks = []
    for r in range(P):
        c_l = np.ascontiguousarray(part.scatter(coef, r))
        b_l = np.ascontiguousarray(part.scatter(rhs, r))
        ranks.append(CudaSlabOps(kt.Field(c_l.shape, np.float64, cuda).from_numpy(c_l), kt.Field(b_l.shape, np.float64, cuda).from_numpy(b_l),
                                 kt.Field(b_l.shape, np.float64, cuda), part, r))

    def exchange_p():
        for r in range(P - 1):
            lo, up = ranks[r].planes()['p'], ranks[r + 1].planes()['p']
            _, l1r = part.local_owned(r)
            l0u, _ = part.local_owned(r + 1)
            lo[l1r].copy_(up[l0u])
            up[l0u - 1].copy_(lo[l1r - 1])

    def step(fn, fin=None):
        for o in ranks:
            getattr(o, fn)()
        if fin is not None:
            tot = sum(o.dots for o in ranks)
            for o in ranks:
                o.dots.copy_(tot)
                o.scalar(fin)

    for o in ranks:
        o.begin(eps, rhs.size)
    step('cg_init', 5); exchange_p()
    for it in range(rhs.size):
        step('cg_ap_dot', 6)
        step('cg_xr_dot', 7)
        step('cg_update_p'); exchange_p()
        polls = [o.poll() for o in ranks]
        assert len({(p[0], p[1], p[2]) for p in polls}) == 1
        if polls[0][0]:
            break
    done, iters, status, rel = polls[0]
    assert status == st_ref == 0 and abs(iters - it_ref) <= 1
    x = np.concatenate([part.owned_of_local(o.x.to_numpy(), r) for r, o in enumerate(ranks)], axis=0)
    assert rel_l2(x, x_ref) < 1e-8


def _nccl_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import karman_taichi_b200 as kt
        from karman_taichi_b200.slab import CudaSlabOps, DistBiCGSTAB, HaloExchanger, SlabPartition
        coef, rhs = krylov.poisson_sin_system(256, 0.1)
        part = SlabPartition(rhs.shape[0], world)
        c_l = np.ascontiguousarray(part.scatter(coef, rank))
        b_l = np.ascontiguousarray(part.scatter(rhs, rank))
        fc = kt.Field(c_l.shape, np.float64, dev).from_numpy(c_l)
        fb = kt.Field(b_l.shape, np.float64, dev).from_numpy(b_l)
        fx = kt.Field(b_l.shape, np.float64, dev)
        ops = CudaSlabOps(fc, fb, fx, part, rank)
        drv = DistBiCGSTAB(ops, HaloExchanger(part, rank), poll_every=4)
        iters, status = drv.solve(eps=1e-10, max_iter=rhs.size)
        np.save(os.path.join(out_dir, f'x_{rank}.npy'), part.owned_of_local(fx.to_numpy(), rank))
        np.save(os.path.join(out_dir, f'meta_{rank}.npy'), np.array([iters, status]))
        # the same solve with the exchange / all-reduce done by the library over NVLink peer memory (no NCCL in the loop)
        from karman_taichi_b200.slab import P2PSlabSolver
        fx2 = kt.Field(b_l.shape, np.float64, dev)
        p2p = P2PSlabSolver(fc, fb, fx2, part, rank)
        for rep in range(2):                      # twice: the sequence numbers must survive a second solve
            it2, st2 = p2p.solve(eps=1e-10, max_iter=rhs.size)
        np.save(os.path.join(out_dir, f'xp_{rank}.npy'), part.owned_of_local(fx2.to_numpy(), rank))
        np.save(os.path.join(out_dir, f'metap_{rank}.npy'), np.array([it2, st2]))
        # CG over the same slabs: NCCL driver, then the peer-memory kernel path
        from karman_taichi_b200.slab import DistCG
        itc, stc = DistCG(ops, HaloExchanger(part, rank), poll_every=4).solve(eps=1e-10, max_iter=rhs.size)
        np.save(os.path.join(out_dir, f'xc_{rank}.npy'), part.owned_of_local(fx.to_numpy(), rank))
        itc2, stc2 = p2p.solve_cg(eps=1e-10, max_iter=rhs.size)
        np.save(os.path.join(out_dir, f'xcp_{rank}.npy'), part.owned_of_local(fx2.to_numpy(), rank))
        np.save(os.path.join(out_dir, f'metac_{rank}.npy'), np.array([itc, stc, itc2, stc2]))
    finally:
        dist.destroy_process_group()


def test_two_gpu_nccl(cuda, tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs >= 2 GPUs')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    coef, rhs = krylov.poisson_sin_system(256, 0.1)
    x_ref, it_ref, _ = krylov.bicgstab_solve(coef, rhs, eps=1e-10)
    x = np.concatenate([np.load(tmp_path / f'x_{r}.npy') for r in range(world)], axis=0)
    metas = [np.load(tmp_path / f'meta_{r}.npy') for r in range(world)]
    assert all(int(m[1]) == 0 for m in metas) and len({int(m[0]) for m in metas}) == 1
    assert abs(int(metas[0][0]) - it_ref) <= max(3, round(0.1 * it_ref))    # NCCL's summation order is its own
    assert rel_l2(x, x_ref) < 1e-8
    xp = np.concatenate([np.load(tmp_path / f'xp_{r}.npy') for r in range(world)], axis=0)
    mp_ = [np.load(tmp_path / f'metap_{r}.npy') for r in range(world)]
    assert all(int(m[1]) == 0 for m in mp_) and len({int(m[0]) for m in mp_}) == 1
    # peer-memory path: the ranks' local sums are added in rank order (k_p2p_reduce) -> the oracle adding in that order must
    # reproduce it bit for bit; the NCCL path above adds in NCCL's own order and is held to the pairwise-order band
    from karman_taichi_b200.slab import SlabPartition
    from oracle.device_order import SlabOrder, installed
    with installed(SlabOrder(rhs.shape, SlabPartition(rhs.shape[0], world).bounds,
                             num_sms=torch.cuda.get_device_properties(cuda).multi_processor_count)):
        x_dev, it_dev, _ = krylov.bicgstab_solve(coef, rhs, eps=1e-10)
    assert int(mp_[0][0]) == it_dev
    np.testing.assert_array_equal(xp, x_dev)
    assert rel_l2(xp, x_ref) < 1e-8
    xc_ref, itc_ref, _ = krylov.cg_solve(coef, rhs, eps=1e-10)
    mc = [np.load(tmp_path / f'metac_{r}.npy') for r in range(world)]
    assert all(int(m[1]) == 0 and int(m[3]) == 0 for m in mc) and len({(int(m[0]), int(m[2])) for m in mc}) == 1
    assert abs(int(mc[0][0]) - itc_ref) <= 1 and abs(int(mc[0][2]) - itc_ref) <= 1
    for tag in ('xc', 'xcp'):
        xc = np.concatenate([np.load(tmp_path / f'{tag}_{r}.npy') for r in range(world)], axis=0)
        assert rel_l2(xc, xc_ref) < 1e-8


KARMAN_MP = dict(nx=96, ny=48, cyl=(64, 32, 144))


def _slab_karman_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import karman_taichi_b200 as kt
        s = kt.SlabKarmanSolver(rank, world, device=dev, peers='ipc', **KARMAN_MP)
        s.init()
        out = []
        for k in range(2):
            out.append((s.sub_iteration(1e-8), dict(s.last_iters), dict(s.last_status)))
        np.savez(os.path.join(out_dir, f'karman_{rank}.npz'), pcor_max=np.array([o[0] for o in out]),
                 iters=np.array([[o[1][w] for w in 'uvp'] for o in out]), status=np.array([[o[2][w] for w in 'uvp'] for o in out]),
                 **{name: s.owned(getattr(s, name)) for name in ('u', 'v', 'p', 'pcor', 'udiv')})
    finally:
        dist.destroy_process_group()


def test_multi_gpu_slab_karman(cuda, tmp_path):
    """SlabKarmanSolver, one process per GPU, peer mappings through CUDA IPC (the deployment mode; the in-process variant is
    tests/test_gpu_slab_threads.py): two sub-iterations of the channel, every field and every Krylov count bit for bit equal
    to the oracle adding in the slabs' order."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs >= 2 GPUs')
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    mp.spawn(_slab_karman_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from karman_taichi_b200.slab import SlabPartition
    from oracle.device_order import SlabAutoOrder, installed
    from oracle.karman import KarmanOracle
    nx = KARMAN_MP['nx']
    part_p = SlabPartition(nx + 2, world)
    part_u = SlabPartition.from_bounds(part_p.bounds[:-1] + [nx + 3])
    o = KarmanOracle(**KARMAN_MP)
    o.init()
    res = [np.load(tmp_path / f'karman_{r}.npz') for r in range(world)]
    with installed(SlabAutoOrder({nx + 3: part_u.bounds, nx + 2: part_p.bounds},
                                 num_sms=torch.cuda.get_device_properties(cuda).multi_processor_count)):
        for k in range(2):
            po = o.sub_iteration(1e-8)
            for r in res:
                assert r['pcor_max'][k] == po
                assert [int(v) for v in r['iters'][k]] == [o.last_iters[w] for w in 'uvp']
                assert [int(v) for v in r['status'][k]] == [o.last_status[w] for w in 'uvp']
    for name in ('u', 'v', 'p', 'pcor', 'udiv'):
        np.testing.assert_array_equal(np.concatenate([r[name] for r in res], axis=0), getattr(o, name), err_msg=name)


def test_p2p_single_rank(cuda):
    """nranks = 1 through kt_p2p_bicgstab_solve: the peer-memory reduce kernel (store, flag, wait, rank-ordered sum, scalar
    algebra) on its own block; must reproduce the plain solver."""
    import karman_taichi_b200 as kt
    from karman_taichi_b200.slab import P2PSlabSolver, SlabPartition
    coef, rhs = krylov.random_dd_system(130, 67, seed=1)
    x_ref, it_ref, _ = krylov.bicgstab_solve(coef, rhs, eps=1e-10)
    fc = kt.Field(coef.shape, np.float64, cuda).from_numpy(coef)
    fb = kt.Field(rhs.shape, np.float64, cuda).from_numpy(rhs)
    fx = kt.Field(rhs.shape, np.float64, cuda)
    s = P2PSlabSolver(fc, fb, fx, SlabPartition(rhs.shape[0], 1), 0)
    for _ in range(2):
        it, st = s.solve(eps=1e-10, max_iter=rhs.size)
    import torch
    from oracle.device_order import SlabOrder, installed
    with installed(SlabOrder(rhs.shape, [0, rhs.shape[0]], num_sms=torch.cuda.get_device_properties(cuda).multi_processor_count)):
        x_dev, it_dev, st_dev = krylov.bicgstab_solve(coef, rhs, eps=1e-10)
    assert (it, st) == (it_dev, st_dev)                      # order pinned: equal
    np.testing.assert_array_equal(fx.to_numpy(), x_dev)
    assert rel_l2(fx.to_numpy(), x_ref) < 1e-8
    ref = kt.BICGSolver(fc, fb, kt.Field(rhs.shape, np.float64, cuda))
    ref.cluster = False
    ref.solve(eps=1e-10)
    assert ref.iterations == it          # identical kernels, identical sums: identical count
    # CG through kt_p2p_cg_solve
    coef, rhs = krylov.poisson_sin_system(150, 0.1)
    xc_ref, itc_ref, _ = krylov.cg_solve(coef, rhs, eps=1e-10)
    fc = kt.Field(coef.shape, np.float64, cuda).from_numpy(coef)
    fb = kt.Field(rhs.shape, np.float64, cuda).from_numpy(rhs)
    fx = kt.Field(rhs.shape, np.float64, cuda)
    s = P2PSlabSolver(fc, fb, fx, SlabPartition(rhs.shape[0], 1), 0)
    it, st = s.solve_cg(eps=1e-10, max_iter=rhs.size)
    assert st == 0 and abs(it - itc_ref) <= 1
    assert rel_l2(fx.to_numpy(), xc_ref) < 1e-8
