"""Where does a cluster-resident BiCGSTAB iteration spend its time?  (development tool)

Builds a -DKT_CLUSTER_PROF variant of the library (clock64() stamps in thread 0 of CTA 0, csrc/kt_cluster.cu) and prints
the ticks per segment and per iteration for a few grid sizes.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, 'gpurun_out', 'libkarman_b200_clprof.so')
os.environ['KT_LIB_PATH'] = out          # read when the package is first imported
os.makedirs(os.path.dirname(out), exist_ok=True)
from karman_taichi_b200 import build as kb  # noqa: E402
kb.build(extra_flags=['-DKT_CLUSTER_PROF'], out=out)

import torch  # noqa: E402
import karman_taichi_b200 as kt  # noqa: E402
from karman_taichi_b200 import _lib  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from microbench import poisson_on_device  # noqa: E402

lib = _lib.load()
prof = lib.kt_debug_cluster_prof
prof.argtypes = [C.POINTER(C.c_longlong), C.c_int]
NAMES = ['P1 loop', 'P1 cl.sync', 'P2 loop (SpMV)', 'sum: warp tree + smem + syncthreads', 'sum: tree16 + DSMEM push', 'sum: cl.sync',
         'sum: final tree16', 'scalar algebra', 'P3 loop (SpMV on s)', 'P4 loop']
dev = torch.device('cuda', 0)
for n in [int(a) for a in sys.argv[1:]] or [48, 128, 176]:
    coef, b, x = poisson_on_device(n, 0.0, np.float64, dev)
    s = kt.BICGSolver(coef, b, x)
    s.max_steps = 60
    s.solve(eps=0.0)
    buf = (C.c_longlong * 16)()
    prof(buf, 1)
    s.solve(eps=0.0)
    torch.cuda.synchronize()
    prof(buf, 0)
    it = s.iterations
    tot = sum(buf[k] for k in range(10))
    print(f'n = {n}: {it} iterations, {tot / it:.0f} ticks / iteration')
    for k, nm in enumerate(NAMES):
        print(f'    {nm:40s} {buf[k] / it:8.0f} ticks / iteration  ({100.0 * buf[k] / tot:4.1f} %)')
