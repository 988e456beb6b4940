#!/usr/bin/env python
"""bench.py -- the contract benchmark of the hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--iters I]

Metric (BASELINE.json): BiCGSTAB Mcell-iters/s = cells x iterations / second / 1e6 on the pressure-correction-type
system S1 (5-point Poisson stencil arrays with a zero-coefficient ghost ring, sin*sin right-hand side,
offset 0), f64.  A *step* is one ``BICGSolver.solve`` of a fixed ``--iters`` BiCGSTAB iterations (eps = 0 so the
count is exact) on synthetic inputs.  Default workload at every N: ``pcor32768`` (config 5, the largest configuration;
it fits one B200: 103 GB).  N > 1 splits the same grid into x-slabs (strong scaling), one rank per GPU under torchrun,
halo rows over NCCL send/recv and the inner products over NCCL all-reduce.

``--workload cavity128 | karman320 | karman640 | karman16384`` measures the other half of BASELINE.json's metric instead, SIMPLE
outer-iterations/s on configs 2, 1 and 4 (one step = one outer iteration through SIMPLESolver / KarmanSolver).

One JSON line on stdout (rank 0): value (inputs resident in HBM), e2e (host buffers in the reference layout through
the C ABI, H2D/D2H inside the timed region), roofline (dominant kernel, timed live with CUDA events), cpu_baseline
(the oracle's C/OpenMP port of the reference algorithm on this box's host cores), clocks, gpu_launches.

``--impl reference`` times the reference's own CPU implementation of the path.  The reference is Python + Taichi 0.9
(``ti.init(arch=ti.cpu)``); Taichi cannot be installed offline for this Python (no wheel, no network), so the arm runs
the C/OpenMP port in ``oracle/c`` that keeps the reference's kernel granularity one to one (kind "port").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n cells per side, description)
    'pcor32768': (32768, 'BASELINE config 5: 32768x32768 pressure-correction BiCGSTAB (x-slabs when N > 1)'),
    'pcor16384': (16384, 'north_star single-B200 target: 16384x16384 pressure-correction BiCGSTAB'),
    'pcor4096': (4096, 'BASELINE config 3: standalone 4096x4096 pressure-correction BiCGSTAB'),
    'pcor1024': (1024, 'development size'),
}
SIMPLE_WORKLOADS = {
    # name: (case, nx, ny, Krylov cap per solve (0 = none), description) -- the other half of BASELINE.json's metric
    'cavity128': ('cavity', 128, 128, 0, 'BASELINE config 2: lid-driven cavity 128x128, Re=100, SIMPLE (bicgsimp.py), alpha_p = 0.02 '
                                         '(the committed 0.1 diverges at 128^2: SURVEY.md 8d)'),
    'karman320': ('karman', 320, 64, 0, 'BASELINE config 1: 320x64 channel + square-cell cylinder, SIMPLE sub-iterations (karman_taichi.py)'),
    'karman640': ('karman', 640, 128, 0, 'the size karman_taichi.py is committed with (nx = 640, ny = 128: karman_taichi.py:11-12), SIMPLE sub-iterations'),
    'karman16384': ('karman', 16384, 3276, 200, 'BASELINE config 4: channel scaled to 16384x3276, full SIMPLE sub-iterations, '
                                                'BiCGSTAB capped at 200 iterations per solve'),
}
SIMPLE_METRIC = 'simple_outer_iters_per_s'
SIMPLE_UNIT = 'outer-iters/s'
METRIC = 'bicgstab_mcell_iters_per_s'
UNIT = 'Mcell-iters/s'
BYTES_PER_CELL_ITER = 28 * 8            # BiCGSTAB algorithmic traffic, f64
# dram__bytes_read.sum + dram__bytes_write.sum per launch and per plane word of each kernel come from the committed
# ncu captures (profiles/r01_launch_list_default.md: the default command, 32770 x 32784-word planes).
FUSED_TRAFFIC_BYTES_PER_WORD = 86.9
NO_S_K3_TRAFFIC = 65.1      # profiles/r01_launch_list_default.md
NO_S_K4_TRAFFIC = 63.9


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpus):
        self.gpus, self.proc, self.lines = set(gpus), None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9 or not f[0].isdigit() or int(f[0]) not in self.gpus:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------
# synthetic S1 system
# ---------------------------------------------------------------------------------------------------------
def s1_rows_numpy(n, g0, g1, offset=0.0):
    """Rows [g0, g1) of the global S1 system in the REFERENCE layout: coef (rows, n+2, 5) AoS, b (rows, n+2).
    Same arrays as oracle.krylov.poisson_sin_system (restated here so the product path never imports oracle/)."""
    NY = n + 2
    rows = g1 - g0
    coef = np.zeros((rows, NY, 5))
    b = np.zeros((rows, NY))
    g = np.arange(g0, g1)
    inner = (g >= 1) & (g <= n)
    coef[inner, 1:n + 1, 0] = 4.0 + offset
    coef[inner, 1:n + 1, 1:5] = -1.0
    coef[g == 1, :, 1] = 0.0
    coef[g == n, :, 2] = 0.0
    coef[:, n, 3] = 0.0
    coef[:, 1, 4] = 0.0
    sj = np.sin(2.0 * np.pi * np.arange(n) / n)
    si = np.where(inner, np.sin(2.0 * np.pi * (g - 1) / n), 0.0)
    b[:, 1:n + 1] = si[:, None] * sj[None, :]
    return coef, b


def s1_rows_device(kt, torch, n, g0, g1, dev, offset=0.0, np_dtype=np.float64):
    """The same rows assembled directly in HBM (value leg: inputs resident before the timed region)."""
    NY = n + 2
    rows = g1 - g0
    coef = kt.Field((rows, NY, 5), np_dtype, dev)
    b = kt.Field((rows, NY), np_dtype, dev)
    x = kt.Field((rows, NY), np_dtype, dev)
    g = torch.arange(g0, g1, device=dev)
    inner = ((g >= 1) & (g <= n)).to(torch.float64)[:, None]
    d = coef.data
    d[0, :, 1:n + 1] = (4.0 + offset) * inner
    for k in range(1, 5):
        d[k, :, 1:n + 1] = -1.0 * inner
    d[1][(g == 1).nonzero().flatten(), :] = 0.0
    d[2][(g == n).nonzero().flatten(), :] = 0.0
    d[3, :, n] = 0.0
    d[4, :, 1] = 0.0
    sj = torch.sin(2.0 * np.pi * torch.arange(n, dtype=torch.float64, device=dev) / n)
    si = torch.sin(2.0 * np.pi * (g.to(torch.float64) - 1.0) / n) * inner[:, 0]
    b.data[:, 1:n + 1] = (si[:, None] * sj[None, :]).to(b.data.dtype)
    return coef, b, x


# ---------------------------------------------------------------------------------------------------------
# CPU arm (oracle C/OpenMP port) -- cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------------------
def cpu_port_run(n_sample, iters, reps):
    """Time `reps` fixed-iteration BiCGSTAB runs of the C/OpenMP port on an n_sample^2 S1 system.  Returns
    (best seconds per run, threads)."""
    from oracle import cport
    # every host core, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which would turn the reference arm
    # into a single-thread run at N > 1
    threads = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    cport.load()
    cport.set_threads(threads)
    coef, b = s1_rows_numpy(n_sample, 0, n_sample + 2)
    work = np.zeros(9 * b.size)
    best = float('inf')
    for _ in range(reps):
        t = time.perf_counter()
        _, it, st, _ = cport.bicgstab_solve(coef, b, eps=0.0, max_iter=iters, work=work)
        best = min(best, time.perf_counter() - t)
        assert it == iters, (it, st)
    return best, threads


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    n, desc = WORKLOADS[a.workload]
    n_s, its = min(n, 4096), 6
    cpu_port_run(n_s, 2, 1)                       # page-in / first touch
    for _ in range(max(a.warmup, 1) - 1):
        cpu_port_run(n_s, its, 1)
    t0 = time.perf_counter()
    per = [cpu_port_run(n_s, its, 1)[0] for _ in range(a.steps)]
    total = time.perf_counter() - t0
    threads = cpu_port_run(n_s, 1, 1)[1]
    val = n_s * n_s * its * a.steps / sum(per) / 1e6
    sample = f'{n_s}x{n_s} sub-grid of the S1 system, {its} BiCGSTAB iterations per step (reference kernel granularity: 14 loops + 5 reductions per iteration)'
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': 1e3 * sum(per) / a.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': {'workload': a.workload, 'description': desc, 'cpu_sample': sample},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'note': 'Taichi ti.cpu itself cannot run offline (no taichi wheel for this Python); C/OpenMP port of cgsolver.py/bicgsolver.py',
            'wall_s': total}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if a.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun exactly as the driver does
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={a.gpus}', '--master-addr', '127.0.0.1',
               '--master-port', str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    assert world == a.gpus, f'--gpus {a.gpus} but WORLD_SIZE={world}'
    assert torch.cuda.is_available(), 'bench.py needs CUDA (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import karman_taichi_b200 as kt
    from karman_taichi_b200 import _lib
    from karman_taichi_b200.cgsolver import _p
    from karman_taichi_b200.slab import CudaSlabOps, DistBiCGSTAB, HaloExchanger, SlabPartition
    lib = _lib.load()

    n, desc = WORKLOADS[a.workload]
    NX = NY = n + 2
    part = SlabPartition(NX, world)
    g0, g1 = part.stored(rank)
    l0, l1 = part.local_owned(rank)
    np_dtype = np.float64 if a.dtype == 'f64' else np.float32
    w_bytes = 8 if a.dtype == 'f64' else 4
    coef, b, x = s1_rows_device(kt, torch, n, g0, g1, dev, np_dtype=np_dtype)
    cells = n * n
    iters = a.iters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if world == 1:
        solver = kt.BICGSolver(coef, b, x)
        solver.max_steps = iters
        handle = solver._hd

        def step():
            solver.solve(eps=0.0, quiet=True)          # the public call: BICGSolver.solve (bicgsolver.py:88)
            assert solver.iterations == iters, (solver.iterations, solver.status)
    elif a.comm == 'nccl':
        ops = CudaSlabOps(coef, b, x, part, rank)
        drv = DistBiCGSTAB(ops, HaloExchanger(part, rank))
        handle = ops._hd

        def step():
            it, st = drv.solve(eps=0.0, max_iter=iters, fixed_iters=True)
            assert it == iters, (it, st)
    else:
        from karman_taichi_b200.slab import P2PSlabSolver
        try:
            drv = P2PSlabSolver(coef, b, x, part, rank)
            handle = drv._hd
        except _lib.KtError as e:        # raised on every rank together (slab.py): e.g. no peer access on this box
            if rank == 0:
                print(f'bench: peer-memory setup failed ({e}); using the NCCL exchange', file=sys.stderr, flush=True)
            a.comm = 'nccl'
            ops = CudaSlabOps(coef, b, x, part, rank)
            drv = DistBiCGSTAB(ops, HaloExchanger(part, rank))
            handle = ops._hd

        def step():
            it, st = drv.solve(eps=0.0, max_iter=iters, fixed_iters=True)
            assert it == iters, (it, st)

    # ---- value leg: inputs resident in HBM -------------------------------------------------------------
    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(range(world))
    if rank == 0:
        sampler.start()
    _lib.check(lib.kt_prof_begin(handle.h, 4 * iters * a.steps + 64))
    launches0 = handle.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms_total = allmax(e0.elapsed_time(e1))
    launches = int(allsum(handle.launch_count() - launches0))
    tot = (C.c_double * 8)()
    cnt = (C.c_long * 8)()
    _lib.check(lib.kt_prof_end(handle.h, tot, cnt, 8))
    clocks = sampler.stop() if rank == 0 else None
    value = cells * iters * a.steps / (ms_total * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (largest share of the step), timed live with CUDA-event pairs ---------------
    peak, peak_kind = peaks()
    fused = cnt[0] == 0                       # BICGSolver.fuse_p_update: K1 folded into K2 (11 N instead of 4 N + 8 N)
    store_s = world > 1 and a.comm == 'nccl'  # the single-GPU and peer-memory paths recompute s in K4 instead of storing it
    classes = {0: ('p_update', 'k_bicg_p: p = r + beta (p - omega v)', 4, 31.6),
               1: (('p_update+v=Ap+dot', 'k_stencil<SRC_P,EPI_DOT_Q_ST>: p = r + beta (p - omega v), v = A p, (r~, v)', 11, FUSED_TRAFFIC_BYTES_PER_WORD)
                   if fused else ('v=Ap+dot', 'k_stencil<SRC_X,EPI_DOT_Q>: v = A p, (r~, v)', 8, 64.0)),
               2: (('s,t+2dots', 'k_stencil<SRC_S,EPI_ST>: s = r - alpha v, t = A s, (t,s), (t,t)', 9, 73.1) if store_s else
                   ('t=As+2dots', 'k_stencil<SRC_S,EPI_ST>: t = A (r - alpha v) [s not stored], (t,s), (t,t)', 8, NO_S_K3_TRAFFIC)),
               3: (('x,r+2dots', 'k_bicg_xr: x += alpha p + omega s, r = s - omega t, (r,r), (r,r~)', 7, 55.9) if store_s else
                   ('x,r+2dots', 'k_bicg_xr: s = r - alpha v, x += alpha p + omega s, r = s - omega t, (r,r), (r,r~)', 8, NO_S_K4_TRAFFIC))}
    tots = [allmax(tot[i]) for i in range(4)]
    dom = max(range(4), key=lambda i: tots[i])
    d_name, d_desc, d_words, d_traffic = classes[dom]
    d_ms = tots[dom] / max(cnt[dom], 1)
    words_local = (l1 - l0) * NY
    alg_bytes = d_words * w_bytes * words_local
    achieved = alg_bytes / (d_ms * 1e-3) / 1e9
    kernel_share = {classes[i][0]: tots[i] / ms_total for i in range(4) if cnt[i]}
    iter_bytes = (27 if fused else 28) * w_bytes
    cell_iter_bytes = 28 * w_bytes        # 224 B f64, 112 B f32
    roofline = {'bound': 'hbm', 'kernel': f'{d_desc}  [{d_words} N words]',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_kind': peak_kind,
                'traffic': (d_traffic * (l1 - l0) * coef.pitch) if (d_traffic and a.dtype == 'f64') else None,
                'traffic_source': 'profiles/r01_launch_list_default.md (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of this command), scaled by plane words',
                'ms_per_launch': d_ms, 'launches_timed': int(cnt[dom]), 'algorithmic_bytes_per_launch': alg_bytes,
                'kernel_share_of_step': kernel_share,
                'whole_iteration': {'algorithmic_bytes_per_cell_iter': cell_iter_bytes, 'bytes_per_cell_iter_as_run': iter_bytes,
                                    'achieved_gbs_per_gpu': cell_iter_bytes * NX * NY * iters * a.steps / world / (ms_total * 1e-3) / 1e9,
                                    'frac': cell_iter_bytes * NX * NY * iters * a.steps / world / (ms_total * 1e-3) / 1e9 / peak}}

    # ---- e2e leg: host buffers in the reference layout, H2D + solve + D2H inside the timed region --------------------
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, kt, _lib, lib, torch, dist, world, rank, dev, n, g0, g1, coef, b, x, step, handle, cells, iters, barrier, allmax, allsum)

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n_s, its = 4096, 6
        cpu_port_run(n_s, 2, 1)
        t, threads = cpu_port_run(n_s, its, 3)
        cpu = {'value': n_s * n_s * its / t / 1e6, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': f'{n_s}x{n_s} sub-grid of the same S1 system, {its} BiCGSTAB iterations, best of 3 (C/OpenMP port of bicgsolver.py, '
                         'reference kernel granularity)'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
                'ms_per_step': ms_total / a.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': a.dtype,
                'data': 'synthetic',
                'config': {'workload': a.workload, 'description': desc, 'grid': [n, n], 'bicgstab_iters_per_step': iters,
                           'system': 'S1 Poisson-sin, offset 0, zero-coefficient ghost ring (SURVEY.md 8d)',
                           'parallelism': (f'x-slab x{world}, halo + all-reduce: ' + ('NVLink peer memory inside the library' if a.comm == 'p2p' else 'NCCL send/recv + all_reduce')) if world > 1 else 'single GPU',
                           'cache': f'working set {12 * NX * NY * w_bytes / 1e9:.1f} GB vs 126 MB L2 (inputs larger than L2: no flush needed)'},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(a, kt, _lib, lib, torch, dist, world, rank, dev, n, g0, g1, coef, b, x, step, handle, cells, iters, barrier, allmax, allsum):
    """The reference-facing call with HOST buffers: coef[i][j][k] (AoS) and b in pinned host memory are copied to the
    device, converted to the kernel layout, solved, and x is copied back -- every step, all inside the timed region."""
    rows, NY = g1 - g0, n + 2
    w_bytes = 8 if a.dtype == 'f64' else 4
    t_dtype = torch.float64 if a.dtype == 'f64' else torch.float32
    need = rows * NY * w_bytes * 7
    try:
        avail = int([l for l in open('/proc/meminfo') if l.startswith('MemAvailable')][0].split()[1]) * 1024
    except Exception:
        avail = None
    if avail is not None and need * (world if world > 1 else 1) > 0.8 * avail:
        return {'value': None, 'unit': UNIT, 'skipped': f'host needs {need / 1e9:.0f} GB pinned per rank, {avail / 1e9:.0f} GB available'}
    try:
        coef_h = torch.empty((rows, NY, 5), dtype=t_dtype, pin_memory=True)
        b_h = torch.empty((rows, NY), dtype=t_dtype, pin_memory=True)
        x_h = torch.empty((rows, NY), dtype=t_dtype, pin_memory=True)
    except Exception as e:   # pinned allocation refused
        return {'value': None, 'unit': UNIT, 'skipped': f'pinned host allocation failed: {e}'}
    CH = 8192                                    # fill in row chunks to bound temporary memory
    for r in range(0, rows, CH):
        c_np, b_np = s1_rows_numpy(n, g0 + r, min(g0 + r + CH, g1))
        coef_h[r:r + CH] = torch.from_numpy(c_np).to(t_dtype)
        b_h[r:r + CH] = torch.from_numpy(b_np).to(t_dtype)
    staging_bytes = 256 << 20
    staging = torch.empty(staging_bytes // 8, dtype=torch.float64, device=dev)
    h = handle.h
    P = lambda t: C.c_void_p(t.data_ptr())

    def e2e_step():
        handle.use_current_stream()
        _lib.check(lib.kt_upload_coef(h, C.c_void_p(coef.ptr), P(coef_h), P(staging), staging_bytes))
        _lib.check(lib.kt_upload_field(h, C.c_void_p(b.ptr), P(b_h)))
        step()
        _lib.check(lib.kt_download_field(h, P(x_h), C.c_void_p(x.ptr)))

    k_e = max(1, min(a.steps, 3))
    e2e_step()                                   # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k_e):
        e2e_step()
    e1.record()
    barrier()
    ms = allmax(e0.elapsed_time(e1))
    h2d = int(allsum(float((coef_h.numel() + b_h.numel()) * w_bytes)))     # bytes of the tensors copied, all ranks
    d2h = int(allsum(float(x_h.numel() * w_bytes)))
    return {'value': cells * iters * k_e / (ms * 1e-3) / 1e6, 'unit': UNIT,
            'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
            'steps': k_e, 'ms_per_step': ms / k_e,
            'call': 'kt_upload_coef + kt_upload_field + BICGSolver.solve + kt_download_field (= kt_bicgstab_solve_host) on pinned host buffers'}


# ---------------------------------------------------------------------------------------------------------
# SIMPLE outer-iterations / s (BASELINE configs 1, 2, 4): --workload cavity128 | karman320 | karman16384
# ---------------------------------------------------------------------------------------------------------
def simple_oracle_rate(case, nx, ny, cap, steps):
    """The NumPy oracle (oracle/simple.py, oracle/karman.py) timed on the host: outer iterations / s on a bounded
    sample.  Returns (rate scaled to the workload's grid, description)."""
    cells = nx * ny
    if case == 'cavity':
        from oracle.simple import SimpleOracle
        o = SimpleOracle(1.0, 1.0, nx, ny)
        o.bc['n'][1] = 1.0
        o.alpha_p = 0.02
        o.init()
        o.step()
        t = time.perf_counter()
        for _ in range(steps):
            o.step()
        dt = time.perf_counter() - t
        return steps / dt, f'{steps} outer iterations of the same {nx}x{ny} cavity, NumPy restatement of bicgsimp.py (1 thread)'
    from oracle.karman import KarmanOracle
    sx, sy = (nx, ny) if cells <= 1 << 20 else (1024, 204)
    cyl = (64, 32, 144) if (sx, sy) == (320, 64) else (sx // 5, sy // 2, int(round((0.1875 * sy) ** 2)))
    o = KarmanOracle(nx=sx, ny=sy, cyl=cyl, dt=1e-5 * 320 / sx)
    if cap:
        o.max_iter = cap
    o.init()
    o.sub_iteration(1e-8)
    t = time.perf_counter()
    for _ in range(steps):
        o.sub_iteration(1e-8)
    dt = time.perf_counter() - t
    scale = (sx * sy) / cells
    what = f'{steps} sub-iterations of the {sx}x{sy} channel' + ('' if scale == 1 else f' (same geometry and Krylov cap; rate scaled by the cell ratio {scale:.3e})')
    return steps / dt * scale, what + ', NumPy restatement of karman_taichi.py (1 thread)'


def run_simple_reference(a):
    if int(os.environ.get('RANK', '0')) != 0:
        return 0
    case, nx, ny, cap, desc = SIMPLE_WORKLOADS[a.workload]
    k = max(1, min(a.steps, 5))
    val, sample = simple_oracle_rate(case, nx, ny, cap, k)
    line = {'impl': 'reference', 'metric': SIMPLE_METRIC, 'value': val, 'unit': SIMPLE_UNIT, 'n_gpus': a.gpus, 'steps': k, 'warmup': 1,
            'ms_per_step': 1e3 / val, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': a.workload, 'description': desc, 'cpu_sample': sample},
            'cpu_baseline': {'value': val, 'unit': SIMPLE_UNIT, 'cores': 1, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': SIMPLE_UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'note': 'Taichi ti.cpu itself cannot run offline; NumPy restatement of the reference (oracle/)'}
    print(json.dumps(line), flush=True)
    return 0


def run_simple(a):
    """One step = one SIMPLE outer iteration (bicgsimp.py:313-317) / sub-iteration (karman_taichi.py:633-652) through the
    reference-facing classes.  Not sharded (BASELINE configs 1, 2, 4 are single-GPU): N > 1 runs N replicas."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if a.gpus > 1 and world == 1:
        port = 29500 + (os.getpid() % 2000)
        return subprocess.call([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={a.gpus}', '--master-addr',
                                '127.0.0.1', '--master-port', str(port), os.path.abspath(__file__)] + sys.argv[1:])
    assert torch.cuda.is_available(), 'bench.py needs CUDA (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import karman_taichi_b200 as kt
    from karman_taichi_b200 import _lib
    from karman_taichi_b200.cgsolver import _Handle
    lib = _lib.load()
    case, nx, ny, cap, desc = SIMPLE_WORKLOADS[a.workload]
    cells = nx * ny
    if case == 'cavity':
        s = kt.lid_driven_cavity(nx, device=dev)
        s.alpha_p = 0.02
        s.init()
        s._make_solvers()
        solvers = {'u': s.u_momentum_solver, 'v': s.v_momentum_solver, 'p': s.p_correction_solver}
        step = s.step
        its = lambda: s.log[-1][2]
        words = {'u': 28, 'v': 28, 'p': 16}           # BiCGSTAB 28 N, CG 16 N words per iteration
    else:
        cyl = (64, 32, 144) if (nx, ny) == (320, 64) else (nx // 5, ny // 2, int(round((0.1875 * ny) ** 2)))
        s = kt.KarmanSolver(nx=nx, ny=ny, cyl=cyl, dt=1e-5 * 320 / nx, device=dev)
        s.init()
        solvers = {'u': s.u_solver, 'v': s.v_solver, 'p': s.p_solver}
        step = s.sub_iteration
        its = lambda: dict(s.last_iters)
        words = {'u': 31, 'v': 31, 'p': 31}           # Jacobi-preconditioned BiCGSTAB: 31 N words per iteration
    if cap:
        for sol in solvers.values():
            sol.max_steps = cap
    handles = [s._hd] + [sol._hd for sol in solvers.values()]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(range(world))
    if rank == 0:
        sampler.start()
    hp = solvers['p']._hd
    _lib.check(lib.kt_prof_begin(hp.h, 4 * (cap or 4096) * a.steps + 64))
    l0 = sum(h.launch_count() for h in handles)
    tot_its = {'u': 0, 'v': 0, 'p': 0}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
        for k, v in its().items():
            tot_its[k] += v
    e1.record()
    barrier()
    ms_total = allmax(e0.elapsed_time(e1))
    launches = (sum(h.launch_count() for h in handles) - l0) * world
    tot = (C.c_double * 8)()
    cnt = (C.c_long * 8)()
    _lib.check(lib.kt_prof_end(hp.h, tot, cnt, 8))
    clocks = sampler.stop() if rank == 0 else None
    value = a.steps * world / (ms_total * 1e-3)

    # roofline: whole-step algorithmic bytes (~51 N words of assembly / BC / relaxation / correction per outer
    # iteration + the Krylov iterations actually run); the dominant kernel class of the p' solve when it streams from HBM
    peak, peak_kind = peaks()
    N = (nx + 2) * (ny + 2)
    step_bytes = 8.0 * N * (51 * a.steps + sum(words[k] * tot_its[k] for k in words))
    achieved = step_bytes / (ms_total * 1e-3) / 1e9
    streamed = sum(cnt[i] for i in range(4)) > 0
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_kind': peak_kind, 'traffic': None,
                'kernel': 'whole SIMPLE step: algorithmic bytes of assembly + Krylov iterations actually run / step time',
                'krylov_iterations_per_step': {k: v / a.steps for k, v in tot_its.items()}, 'krylov_cap': cap,
                'krylov_mcell_iters_per_s': cells * sum(tot_its.values()) / (ms_total * 1e-3) / 1e6}
    if streamed:
        names = {0: ('p_update', 4), 1: ('v=A M^-1 p + dot', 9), 2: ('t=A M^-1 s + 2 dots', 9), 3: ('x,r + 2 dots', 9)}
        dom = max(range(4), key=lambda i: tot[i])
        d_ms = tot[dom] / max(cnt[dom], 1)
        # ncu dram__bytes_read.sum + dram__bytes_write.sum per plane word of the Jacobi-variant kernels
        # (profiles/r01_launch_list_karman16384.md, working launches only)
        jac_traffic = {0: 30.6, 1: 71.7, 2: 73.3, 3: 70.9}
        roofline['traffic'] = jac_traffic[dom] * N if case == 'karman' else None
        roofline['traffic_source'] = 'profiles/r01_launch_list_karman16384.md, dominant kernel, per launch, scaled by plane words'
        roofline['dominant_kernel'] = {'name': names[dom][0], 'words': names[dom][1], 'ms_per_launch': d_ms,
                                       'achieved': names[dom][1] * 8.0 * N / (d_ms * 1e-3) / 1e9,
                                       'frac': names[dom][1] * 8.0 * N / (d_ms * 1e-3) / 1e9 / peak,
                                       'share_of_step': {names[i][0]: tot[i] / ms_total for i in range(4)}}
    else:
        roofline['note'] = ('the Krylov solves run as ONE resident kernel each -- a thread-block cluster (csrc/kt_cluster.cu) up to ~35 k '
                            'unknowns, a persistent cooperative kernel (csrc/kt_coop.cu) above -- and are bound by the latency of the '
                            'three dependent inner products per iteration, not by HBM')

    # e2e: u, v, p arrive in pinned host arrays and go back every step (kt_upload_field / kt_download_field)
    e2e = None
    if not a.no_e2e:
        fields = [s.u, s.v, s.p]
        hs = [_Handle(f) for f in fields]
        host = [torch.from_numpy(f.to_numpy()).pin_memory() for f in fields]
        P = lambda t: C.c_void_p(t.data_ptr())

        def e2e_step():
            for f, h, t in zip(fields, hs, host):
                h.use_current_stream()
                _lib.check(lib.kt_upload_field(h.h, C.c_void_p(f.ptr), P(t)))
            step()
            for f, h, t in zip(fields, hs, host):
                _lib.check(lib.kt_download_field(h.h, P(t), C.c_void_p(f.ptr)))
        k_e = max(1, min(a.steps, 5))
        e2e_step()
        barrier()
        e0.record()
        for _ in range(k_e):
            e2e_step()
        e1.record()
        barrier()
        ms = allmax(e0.elapsed_time(e1))
        nbytes = sum(t.numel() * 8 for t in host) * world
        e2e = {'value': k_e * world / (ms * 1e-3), 'unit': SIMPLE_UNIT, 'h2d_bytes_per_step': nbytes, 'd2h_bytes_per_step': nbytes,
               'steps': k_e, 'ms_per_step': ms / k_e,
               'call': 'kt_upload_field(u, v, p) + SIMPLESolver.step / KarmanSolver.sub_iteration + kt_download_field(u, v, p) on pinned host buffers'}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, sample = simple_oracle_rate(case, nx, ny, cap, 3)
        cpu = {'value': v, 'unit': SIMPLE_UNIT, 'cores': 1, 'kind': 'port', 'sample': sample}
    if rank == 0:
        line = {'metric': SIMPLE_METRIC, 'value': value, 'unit': SIMPLE_UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
                'ms_per_step': ms_total / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic',
                'config': {'workload': a.workload, 'description': desc, 'grid': [nx, ny],
                           'parallelism': 'single GPU' if world == 1 else f'{world} independent replicas (the SIMPLE configs are not sharded)',
                           'cache': ('working set >> 126 MB L2' if cells > 1 << 22 else
                                     'working set resident in shared memory / L2 by design (reference-sized grid); inputs change every step')},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='pcor32768', choices=sorted(WORKLOADS) + sorted(SIMPLE_WORKLOADS))
    ap.add_argument('--iters', type=int, default=50, help='BiCGSTAB iterations per step')
    ap.add_argument('--dtype', default='f64', choices=['f64', 'f32'],
                    help='pcor workloads: field precision (the reference computes in f64; f32 is the north_star 1e-4 variant)')
    ap.add_argument('--comm', default='p2p', choices=['p2p', 'nccl'],
                    help='N > 1: halo exchange + all-reduce inside the library over NVLink peer memory (p2p) or via NCCL')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == 'ours':
        a.warmup = 3                                   # timing rules: W >= 3
    if a.workload in SIMPLE_WORKLOADS:
        sys.exit(run_simple_reference(a) if a.impl == 'reference' else run_simple(a))
    sys.exit(run_reference(a) if a.impl == 'reference' else run_ours(a))


if __name__ == '__main__':
    main()
