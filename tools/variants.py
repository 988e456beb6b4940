"""Build tuning variants of libkarman_b200.so (development tool): python tools/variants.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from karman_taichi_b200 import build

VARIANTS = {
    'w1_minb4': ['-DKT_STENCIL_W=1', '-DKT_STENCIL_MINB=4'],
    'w1_minb5': ['-DKT_STENCIL_W=1', '-DKT_STENCIL_MINB=5'],
    'w1_minb6': ['-DKT_STENCIL_W=1', '-DKT_STENCIL_MINB=6'],
    'w2': ['-DKT_STENCIL_W=2', '-DKT_STENCIL_MINB=1'],
}
if __name__ == '__main__':
    d = os.path.join(ROOT, 'gpurun_variants')
    os.makedirs(d, exist_ok=True)
    for name, flags in VARIANTS.items():
        out = os.path.join(d, f'libkt_{name}.so')
        build.build(force=True, extra_flags=flags, out=out)
        print(out)
