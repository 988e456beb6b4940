#!/usr/bin/env python
"""Does the SPACING between the planes a kernel streams matter?  (profiles/r02_plane_spacing.md)

The x/r kernel of BiCGSTAB (k_bicg_xr) reads 6 planes and writes 2 at the same element offset.  When every plane is its own
torch allocation the planes sit a (rounded) plane size apart; this probe carves the planes out of ONE arena at a chosen
spacing and times the kernel through its unit entry point (kt_bicg_xr: 7 planes, s stored) and the p-update
(kt_bicg_update_p: 4 planes), rows x 32770 reals per plane.

    python tools/plane_offset_probe.py [--rows 16386] [--reps 20]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import karman_taichi_b200 as kt                      # noqa: E402
from karman_taichi_b200 import _lib                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=16386)
    ap.add_argument('--ny', type=int, default=32770)
    ap.add_argument('--reps', type=int, default=20)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    proto = kt.Field((a.rows, a.ny), np.float64, dev)
    plane = proto.data.numel()                       # elements, pitch included
    pbytes = plane * 8
    from karman_taichi_b200.cgsolver import _Handle
    hd = _Handle(proto)
    MiB = 1 << 20
    rounded = -(-pbytes // (2 * MiB)) * 2 * MiB      # what torch's allocator spaces separate big allocations by
    spacings = [('contiguous (plane bytes)', pbytes), ('2 MiB-rounded (separate torch allocations)', rounded),
                ('rounded + 2 MiB', rounded + 2 * MiB), ('rounded + 4 MiB', rounded + 4 * MiB),
                ('rounded + 1 MiB', rounded + MiB), ('rounded + 256 KiB', rounded + 256 * 1024), ('rounded + 36 KiB', rounded + 36 * 1024),
                ('rounded + 4 KiB + 128 B', rounded + 4096 + 128)]
    nplanes = 7
    arena = torch.zeros((max(s for _, s in spacings) * nplanes) // 8 + 64, dtype=torch.float64, device=dev)
    base = arena.data_ptr()
    base += (-base) % (2 * MiB)                       # 2 MiB-aligned start, like a fresh cudaMalloc
    print(f'plane {pbytes / 2**30:.4f} GiB = {pbytes} B ({pbytes // 256} x 256 B), rounded spacing {rounded // (2 * MiB)} x 2 MiB')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = (C.c_double * 2)()
    P = C.c_void_p
    for name, sp in spacings:
        ptr = [base + k * sp for k in range(nplanes)]
        x, r, p, s, t, b, v = ptr
        res = {}
        for kern in ('xr', 'p'):
            ts = []
            for rep in range(a.reps + 3):
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
        gb_xr, gb_p = 7 * pbytes / 1e9, 4 * pbytes / 1e9
        print(f'{name:48s} spacing {sp:>12d} B  xr {res["xr"]:.3f} ms ({gb_xr / res["xr"] :.2f} TB/s)   p-update {res["p"]:.3f} ms ({gb_p / res["p"]:.2f} TB/s)',
              flush=True)


if __name__ == '__main__':
    main()
