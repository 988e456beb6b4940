"""Kernel-level timing on one GPU (development tool; bench.py is the contract benchmark).

    python tools/microbench.py --n 16384 [--iters 20] [--dtype f64] [--cg]

Builds the S1 Poisson-sin system on the device, times the fixed-iteration BiCGSTAB / CG drivers and the
individual kernels with CUDA events, and prints achieved algorithmic GB/s against MEASURED_PEAKS.json.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import karman_taichi_b200 as kt  # noqa: E402
from karman_taichi_b200 import _lib  # noqa: E402
from karman_taichi_b200.cgsolver import _p  # noqa: E402


def poisson_on_device(n, offset, dtype, dev):
    """S1 system (oracle.krylov.poisson_sin_system) assembled directly in HBM."""
    NX = NY = n + 2
    coef = kt.Field((NX, NY, 5), dtype, dev)
    b = kt.Field((NX, NY), dtype, dev)
    x = kt.Field((NX, NY), dtype, dev)
    d = coef.data
    d[0, 1:n + 1, 1:n + 1] = 4.0 + offset
    d[1:5, 1:n + 1, 1:n + 1] = -1.0
    d[1, 1, :] = 0.0
    d[2, n, :] = 0.0
    d[3, :, n] = 0.0
    d[4, :, 1] = 0.0
    s = torch.sin(2.0 * np.pi * torch.arange(n, dtype=torch.float64, device=dev) * 2.0 / (2 * n))
    b.data[1:n + 1, 1:n + 1] = torch.outer(s, s).to(b.dtype)
    return coef, b, x


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'], 'measured'
    except Exception:
        return 6650.0, 'fallback'


def time_call(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=4096)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--dtype', default='f64')
    ap.add_argument('--kernels', action='store_true', help='also time individual kernels')
    ap.add_argument('--brief', default=None, help='print one line tagged with this label')
    ap.add_argument('--fused', action='store_true', help='BiCGSTAB with the p-update fused into v = A p (27 N words)')
    ap.add_argument('--no-s', action='store_true', help='BiCGSTAB without storing s (K4 recomputes it from r and v)')
    a = ap.parse_args()
    dtype = np.float64 if a.dtype == 'f64' else np.float32
    w = 8 if a.dtype == 'f64' else 4
    dev = torch.device('cuda', 0)
    coef, b, x = poisson_on_device(a.n, 0.1, dtype, dev)
    N = coef.NX * coef.pitch          # words per plane actually streamed (pad columns included)
    cells = a.n * a.n
    peak, how = peaks()
    out = {'n': a.n, 'dtype': a.dtype, 'plane_words': N, 'peak_gbs': peak, 'peak_kind': how}

    s = kt.BICGSolver(coef, b, x)
    L, h = s._lib, s._h
    ws = s._ws(s=(None if a.no_s else s.s.ptr), t=s.t.ptr)
    if a.fused:
        p2, ap2 = kt.Field(b.shape, dtype, dev), kt.Field(b.shape, dtype, dev)
        ws.p2, ws.Ap2 = p2.ptr, ap2.ptr
    ms = C.c_float()

    def run_bicg():
        _lib.check(L.kt_bicgstab_bench(h, _p(coef), _p(b), _p(x), C.byref(ws), a.iters, C.byref(ms)))
        return ms.value

    for _ in range(2):
        run_bicg()
    t = min(run_bicg() for _ in range(5))
    per_it = t / a.iters
    out['bicgstab'] = {'ms_per_iter': per_it, 'mcell_iters_per_s': cells / per_it / 1e3,
                       'alg_gbs': 28 * N * w / per_it / 1e6, 'frac': 28 * N * w / per_it / 1e6 / peak}

    c = kt.CGSolver(coef, b, x)
    wsc = c._ws()

    def run_cg():
        _lib.check(L.kt_cg_bench(c._h, _p(coef), _p(b), _p(x), C.byref(wsc), a.iters, C.byref(ms)))
        return ms.value

    for _ in range(2):
        run_cg()
    t = min(run_cg() for _ in range(5))
    per_it = t / a.iters
    out['cg'] = {'ms_per_iter': per_it, 'mcell_iters_per_s': cells / per_it / 1e3,
                 'alg_gbs': 16 * N * w / per_it / 1e6, 'frac': 16 * N * w / per_it / 1e6 / peak}

    if a.kernels or a.brief is not None:
        d2 = (C.c_double * 2)()
        kern = {
            'copy(2N)': (2, lambda: _lib.check(L.kt_copy(h, _p(s.t), _p(s.s)))),
            'axpy(3N)': (3, lambda: _lib.check(L.kt_axpy(h, _p(s.t), 0.5, _p(s.s)))),
            'bicg_update_p(4N)': (4, lambda: _lib.check(L.kt_bicg_update_p(h, _p(s.p), _p(s.r), _p(s.Ap), 0.5, 0.5))),
            'spmv(7N)': (7, lambda: _lib.check(L.kt_spmv(h, _p(coef), _p(s.p), _p(s.Ap)))),
            'spmv_dot(8N)': (8, lambda: _lib.check(L.kt_spmv_dot(h, _p(coef), _p(s.p), _p(s.Ap), _p(b), d2))),
            'bicg_s_t(9N)': (9, lambda: _lib.check(L.kt_bicg_s_t(h, _p(coef), _p(s.r), _p(s.Ap), 0.5, _p(s.s), _p(s.t), d2))),
            'bicg_xr(7N)': (7, lambda: _lib.check(L.kt_bicg_xr(h, _p(x), _p(s.r), _p(s.p), _p(s.s), _p(s.t), _p(b), 0.1, 0.1, d2))),
            'dot(2N)': (2, lambda: _lib.check(L.kt_dot(h, _p(s.p), _p(s.s), d2))),
            'torch_copy(2N)': (2, lambda: s.t.data.copy_(s.s.data)),
        }
        out['kernels'] = {}
        for name, (words, fn) in kern.items():
            med, mn = time_call(fn)
            out['kernels'][name] = {'ms': med, 'ms_min': mn, 'gbs': words * N * w / med / 1e6,
                                    'frac': words * N * w / med / 1e6 / peak}
    if a.brief is not None:
        print(f"{a.brief:28s} n={a.n} bicg {out['bicgstab']['ms_per_iter']:.4f} ms/it {out['bicgstab']['alg_gbs']:.0f} GB/s "
              f"({out['bicgstab']['frac']:.3f}) | cg {out['cg']['ms_per_iter']:.4f} ms/it {out['cg']['alg_gbs']:.0f} GB/s ({out['cg']['frac']:.3f})"
              + ' | ' + ' '.join(f"{k.split('(')[0]} {v['ms']:.3f}" for k, v in out['kernels'].items()
                                 if k.split('(')[0] in ('bicg_update_p', 'spmv_dot', 'bicg_s_t', 'bicg_xr')))
    else:
        print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
