#!/usr/bin/env python
"""Does the 512-byte phase of an element-wise stream matter?  (profiles/r02_p2p.md)

Rank r > 0 of an x-slab solve owns rows [1, rows) of its local planes, so its element-wise kernels start 1 row = pitch * 8 B
= 262 272 B (at 32768^2) = 128 B past a 512 B boundary into every plane.  This probe times kt_bicg_xr / kt_bicg_update_p on one
GPU with (a) rows [0, n), (b) rows [1, n + 1) and (c) rows [1, n + 1) with every plane pointer moved back by 128 B so that the
first owned element is 512 B-aligned again."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import karman_taichi_b200 as kt                      # noqa: E402
from karman_taichi_b200 import _lib                  # noqa: E402
from karman_taichi_b200.cgsolver import _Handle      # noqa: E402

rows, ny, reps = 16386, 32770, 20
dev = torch.device('cuda', 0)
lib = _lib.load()
fields = [kt.Field((rows + 1, ny), np.float64, dev) for _ in range(7)]
hd = _Handle(fields[0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = (C.c_double * 2)()
P = C.c_void_p
pbytes = rows * fields[0].pitch * 8
for name, r0, shift in (('rows [0, n)', 0, 0), ('rows [1, n+1)', 1, 0), ('rows [1, n+1), pointers - 128 B', 1, -128),
                        ('rows [0, n), pointers + 128 B', 0, 128), ('rows [0, n), pointers + 256 B', 0, 256),
                        ('rows [0, n), pointers + 2048 B', 0, 2048)):
    _lib.check(lib.kt_set_rows(hd.h, r0, r0 + rows - 1))
    x, r, p, s, t, b, v = (f.ptr + 1024 + shift for f in fields)      # +1024: room for the negative shift, keeps 512 B phase
    res = {}
    for kern in ('xr', 'p'):
        ts = []
        for rep in range(reps + 3):
            torch.cuda.synchronize()
            e0.record()
            if kern == 'xr':
                _lib.check(lib.kt_bicg_xr(hd.h, P(x), P(r), P(p), P(s), P(t), P(b), 0.37, 0.83, out))
            else:
                _lib.check(lib.kt_bicg_update_p(hd.h, P(p), P(r), P(v), 0.5, 0.25))
            e1.record()
            torch.cuda.synchronize()
            if rep >= 3:
                ts.append(e0.elapsed_time(e1))
        res[kern] = float(np.median(ts))
    print(f'{name:40s} xr {res["xr"]:.3f} ms ({7 * pbytes / 1e9 / res["xr"]:.2f} TB/s)   p-update {res["p"]:.3f} ms ({4 * pbytes / 1e9 / res["p"]:.2f} TB/s)',
          flush=True)
