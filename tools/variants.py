"""Build tuning variants of libkarman_b200.so (development tool): python tools/variants.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from karman_taichi_b200 import build

VARIANTS = {
    'base': [],
    'minb3': ['-DKT_STENCIL_MINB=3'],
    'minb4': ['-DKT_STENCIL_MINB=4'],
    'su1': ['-DKT_STENCIL_UNROLL=1'],
    'su1_minb4': ['-DKT_STENCIL_UNROLL=1', '-DKT_STENCIL_MINB=4'],
    'su4': ['-DKT_STENCIL_UNROLL=4'],
    'ew2': ['-DKT_EW_UNROLL=2'],
    'ew1': ['-DKT_EW_UNROLL=1'],
    'ew8': ['-DKT_EW_UNROLL=8'],
}
if __name__ == '__main__':
    d = os.path.join(ROOT, 'gpurun_variants')
    os.makedirs(d, exist_ok=True)
    for name, flags in VARIANTS.items():
        out = os.path.join(d, f'libkt_{name}.so')
        build.build(force=True, extra_flags=flags, out=out)
        print(out)
