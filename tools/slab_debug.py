"""Verbose version of tests/test_gpu_slab_threads.py: prints where the in-process multi-rank runs leave the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch
from oracle import krylov
from oracle.device_order import SlabAutoOrder, SlabOrder, installed
from test_gpu_slab_threads import run_ranks, slab_solvers, sms

cuda = torch.device('cuda', 0)


def krylov_case(P, case):
    kind, a, b_ = case
    coef, rhs = krylov.poisson_sin_system(a, b_) if kind == 'poisson' else krylov.random_dd_system(a, b_, seed=0)
    part, solvers = slab_solvers(cuda, coef, rhs, P)
    print(f'--- P={P} case={case} shape={rhs.shape} bounds={part.bounds}', flush=True)
    with installed(SlabOrder(rhs.shape, part.bounds, num_sms=sms(cuda))):
        x_dev, it_dev, st_dev = krylov.bicgstab_solve(coef, rhs, eps=1e-10)
    for k in range(3):
        res = run_ranks(cuda, P, lambda r: solvers[r].solve(eps=1e-10, max_iter=rhs.size))
        x = np.concatenate([part.owned_of_local(s.x.to_numpy(), r) for r, s in enumerate(solvers)], axis=0)
        print(f'  solve {k}: ranks {res} oracle {(it_dev, st_dev)} rel_res {[s.rel_res for s in solvers]} max|x - x_oracle| {np.abs(x - x_dev).max():.3e} '
              f'first bad row {np.argwhere(np.abs(x - x_dev).max(axis=1) > 0)[:3].ravel().tolist()}', flush=True)
    # the host-driven phases (tests/test_gpu_slab.py) on the same system for comparison
    import karman_taichi_b200 as kt
    fx = kt.Field(rhs.shape, np.float64, cuda)
    s1 = kt.BICGSolver(kt.Field(coef.shape, np.float64, cuda).from_numpy(coef), kt.Field(rhs.shape, np.float64, cuda).from_numpy(rhs), fx)
    s1.solve(eps=1e-10, quiet=True)
    print(f'  one GPU: it {s1.iterations} status {s1.status} max|x1 - x_oracle(slab order)| {np.abs(fx.to_numpy() - x_dev).max():.3e}', flush=True)


def karman_case(P, kw, nsub):
    import karman_taichi_b200 as kt
    from oracle.karman import KarmanOracle
    print(f'--- karman P={P} {kw}', flush=True)
    ranks = [kt.SlabKarmanSolver(r, P, device=cuda, peers='local', **kw) for r in range(P)]
    kt.SlabKarmanSolver.connect_local(ranks)
    o = KarmanOracle(**kw)
    one = kt.KarmanSolver(device=cuda, **kw)
    nx, ny = kw['nx'], kw['ny']
    order = SlabAutoOrder({nx + 3: ranks[0].part_u.bounds, nx + 2: ranks[0].part_p.bounds}, num_sms=sms(cuda))
    gather = lambda name: np.concatenate([s.owned(getattr(s, name)) for s in ranks], axis=0)

    def cmp(tag, names, ref):
        for name in names:
            a = gather(name)
            b = getattr(ref, name)
            b = b if isinstance(b, np.ndarray) else b.to_numpy()
            d = np.abs(a - b)
            bad = np.argwhere(d > 0)
            print(f'  {tag:28s} {name:8s} max diff {d.max():.3e} n_bad {len(bad)} first {bad[:4].tolist()}', flush=True)

    run_ranks(cuda, P, lambda r: ranks[r].init())
    o.init(); one.init()
    cmp('init vs oracle', ('u', 'v', 'p'), o)
    with installed(order):
        for k in range(nsub):
            # stage by stage
            run_ranks(cuda, P, lambda r: ranks[r].fill_Au()); one.fill_Au(); o.fill_Au(); o.fill_Av()
            cmp(f'[{k}] fill_Au vs one', ('coef_u', 'b_u', 'M_u', 'Minv_u', 'coef_v', 'b_v', 'M_v', 'Minv_v'), one)
            for w in ('u', 'v'):
                run_ranks(cuda, P, lambda r: ranks[r].bicgstab(w, 1e-8)); one.bicgstab(w, 1e-8)
                o._solve(w, getattr(o, 'coef_' + w), getattr(o, 'M_' + w), getattr(o, 'b_' + w), getattr(o, 'x_' + w), 1e-8, (nx + 1) * ny if w == 'u' else nx * (ny + 1))
                print(f'  [{k}] solve {w}: ranks {[(s.last_iters[w], s.last_status[w]) for s in ranks]} one {(one.last_iters[w], one.last_status[w])} '
                      f'oracle {(o.last_iters[w], o.last_status[w])}', flush=True)
                cmp(f'[{k}] x_{w} vs oracle', (f'x_{w}',), o)
                cmp(f'[{k}] x_{w} vs one', (f'x_{w}',), one)
            run_ranks(cuda, P, lambda r: ranks[r].xu_back()); one.xu_back()
            U = o.u[1:nx + 2, 1:ny + 1]; U[:] = U + o.velo_rel * (o.x_u[1:nx + 2, 1:ny + 1] - U)
            V = o.v[1:nx + 1, 1:ny + 2]; V[:] = V + o.velo_rel * (o.x_v[1:nx + 1, 1:ny + 2] - V)
            cmp(f'[{k}] xu_back vs oracle', ('u', 'v'), o)
            rr = run_ranks(cuda, P, lambda r: ranks[r].correct_conserv()); r1 = one.correct_conserv(); ro = o.correct_conserv()
            print(f'  [{k}] correct_conserv ranks {rr} one {r1} oracle {ro}', flush=True)
            cmp(f'[{k}] conserv vs oracle', ('u', 'udiv'), o)
            run_ranks(cuda, P, lambda r: ranks[r].fill_Ap()); one.fill_Ap(); o.fill_Ap()
            cmp(f'[{k}] fill_Ap vs oracle', ('coef_p', 'b_p'), o)
            cmp(f'[{k}] fill_Ap vs one', ('coef_p', 'b_p', 'M_p', 'Minv_p'), one)
            run_ranks(cuda, P, lambda r: ranks[r].bicgstab('p', 1e-8)); one.bicgstab('p', 1e-8)
            o._solve('p', o.coef_p, o.M_p, o.b_p, o.x_p, 1e-8, nx * ny); o.pcor[1:nx + 1, 1:ny + 1] = o.x_p[1:nx + 1, 1:ny + 1]
            print(f'  [{k}] solve p: ranks {[(s.last_iters["p"], s.last_status["p"]) for s in ranks]} one {(one.last_iters["p"], one.last_status["p"])} '
                  f'oracle {(o.last_iters["p"], o.last_status["p"])}', flush=True)
            cmp(f'[{k}] pcor vs oracle', ('pcor',), o)
            cmp(f'[{k}] pcor vs one', ('pcor',), one)
            mx = run_ranks(cuda, P, lambda r: ranks[r].puv_correction()); m1 = one.puv_correction(); mo = o.puv_correction()
            print(f'  [{k}] puv_correction ranks {mx} one {m1} oracle {mo}', flush=True)
            cmp(f'[{k}] after correction vs oracle', ('u', 'v', 'p'), o)


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    print('CUDA_DEVICE_MAX_CONNECTIONS =', os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS'), 'KT_P2P_TIMEOUT_MS =', os.environ.get('KT_P2P_TIMEOUT_MS'), flush=True)
    if which == 'stress':
        for rep in range(4):
            for P, case in [(2, ('poisson', 96, 0.1)), (4, ('poisson', 64, 0.1)), (5, ('random', 37, 300))]:
                try:
                    krylov_case(P, case)
                except Exception as e:
                    print('  EXC', repr(e), flush=True)
        try:
            karman_case(2, dict(nx=48, ny=16, cyl=(12, 8, 10)), 1)
        except Exception as e:
            print('  EXC', repr(e), flush=True)
    if which in ('all', 'krylov'):
        for P, case in [(2, ('poisson', 96, 0.1)), (3, ('random', 130, 67)), (5, ('random', 37, 300))]:
            try:
                krylov_case(P, case)
            except Exception as e:
                print('  EXC', repr(e), flush=True)
    if which in ('all', 'karman'):
        for P, kw, nsub in [(2, dict(nx=48, ny=16, cyl=(12, 8, 10)), 2), (3, dict(nx=96, ny=48, cyl=(64, 32, 144)), 1)]:
            try:
                karman_case(P, kw, nsub)
            except Exception as e:
                import traceback; traceback.print_exc()
