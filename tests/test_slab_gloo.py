"""N>1 path on CPU: world_size-2 (and 3) gloo runs of the x-slab BiCGSTAB driver with the oracle standing in for the
CUDA phases.  Checks the partition arithmetic, the halo exchange, the dot all-reduce and that the partitioned solve
reproduces the single-domain oracle solve (same iteration count, same solution)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from karman_taichi_b200.slab import DistBiCGSTAB, DistCG, HaloExchanger, SlabPartition  # noqa: E402
from oracle import krylov  # noqa: E402


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_partition_arithmetic():
    for NX, n in [(10, 1), (10, 2), (11, 3), (4098, 8), (8, 8)]:
        p = SlabPartition(NX, n)
        rows = []
        for r in range(n):
            lo, hi = p.owned(r)
            rows += list(range(lo, hi))
            a, b = p.stored(r)
            assert a == max(lo - 1, 0) if r == 0 else a == lo - 1
            assert b == (hi if r == n - 1 else hi + 1)
            l0, l1 = p.local_owned(r)
            assert l1 - l0 == hi - lo and p.local_nx(r) == b - a
            assert a + l0 == lo
        assert rows == list(range(NX))
        sizes = [p.owned(r)[1] - p.owned(r)[0] for r in range(n)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        SlabPartition(3, 4)


def _worker(rank, world, port, case, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from slab_cpu_ops import OracleSlabOps
        kind, a, b, eps = case
        coef, rhs = krylov.poisson_sin_system(a, b) if kind.startswith('poisson') else krylov.random_dd_system(a, b, seed=0)
        part = SlabPartition(rhs.shape[0], world)
        l0, l1 = part.local_owned(rank)
        coef_l = np.ascontiguousarray(part.scatter(coef, rank))
        b_l = np.ascontiguousarray(part.scatter(rhs, rank))
        ops = OracleSlabOps(coef_l, b_l, l0, l1)
        ex = HaloExchanger(part, rank)
        # halo exchange on its own: fill a plane with the global row index, wipe halos, exchange, compare
        a0, _ = part.stored(rank)
        plane = torch.arange(a0, a0 + part.local_nx(rank), dtype=torch.float64)[:, None].repeat(1, rhs.shape[1])
        ref = plane.clone()
        if l0 > 0:
            plane[:l0] = -1.0
        if l1 < plane.shape[0]:
            plane[l1:] = -1.0
        ex.exchange(plane)
        assert torch.equal(plane, ref)
        solver = (DistCG if kind == 'poisson-cg' else DistBiCGSTAB)(ops, ex, poll_every=4)
        iters, status = solver.solve(eps=eps, max_iter=rhs.size)
        x_own = part.owned_of_local(ops.x.numpy(), rank)
        np.save(os.path.join(out_dir, f'x_{rank}.npy'), x_own)
        np.save(os.path.join(out_dir, f'meta_{rank}.npy'), np.array([iters, status]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,case', [(2, ('poisson', 48, 0.1, 1e-10)), (2, ('random', 37, 29, 1e-10)),
                                        (3, ('random', 20, 31, 1e-9)), (2, ('poisson-cg', 40, 0.1, 1e-10)),
                                        (3, ('poisson-cg', 23, 0.0, 1e-9))])
def test_slab_bicgstab_matches_global_oracle(tmp_path, world, case):
    port = free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    kind, a, b, eps = case
    coef, rhs = krylov.poisson_sin_system(a, b) if kind.startswith('poisson') else krylov.random_dd_system(a, b, seed=0)
    x_ref, it_ref, st_ref = (krylov.cg_solve if kind == 'poisson-cg' else krylov.bicgstab_solve)(coef, rhs, eps=eps)
    x = np.concatenate([np.load(tmp_path / f'x_{r}.npy') for r in range(world)], axis=0)
    metas = [np.load(tmp_path / f'meta_{r}.npy') for r in range(world)]
    assert all(int(m[1]) == st_ref == 0 for m in metas)
    assert len({int(m[0]) for m in metas}) == 1          # every rank stopped at the same iteration
    assert abs(int(metas[0][0]) - it_ref) <= 1
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-8
