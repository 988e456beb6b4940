"""CPU-side checks of the drop-in boundary: the shared library builds for sm_100a, loads, and exports every
symbol include/karman_b200.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'karman_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(kt_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ('kt_create', 'kt_destroy', 'kt_spmv', 'kt_spmv_dot', 'kt_dot', 'kt_bicgstab_solve', 'kt_cg_solve',
                 'kt_assemble_uv', 'kt_assemble_p', 'kt_set_bc_cavity', 'kt_relax_uv', 'kt_correct_puv', 'kt_mdiv_max',
                 'kt_last_error', 'kt_bicgstab_solve_host'):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f'libkarman_b200.so does not export: {missing}'


def test_ctypes_binding_covers_header(lib_path):
    from karman_taichi_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.kt_version() >= 100
    assert lib.kt_pitch_for(130, _lib.KT_F64) == 192 == _lib.pitch_for(130, _lib.KT_F64)      # rows padded to 512 bytes
    assert lib.kt_pitch_for(130, _lib.KT_F32) == 256


def test_library_is_sm100a_and_has_no_torch_types(lib_path):
    out = subprocess.run(['cuobjdump', '--list-elf', lib_path], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    nm = subprocess.run(['nm', '-D', '--defined-only', lib_path], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in nm.splitlines() if ' T ' in l]
    assert all('torch' not in s and 'at::' not in s for s in exported)
    assert set(declared_symbols()) <= set(exported)


def test_argument_errors_do_not_need_a_gpu(lib_path):
    from karman_taichi_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.kt_create(0, 8, 8, 7, _lib.KT_F64, None, ctypes.byref(h)) < 0      # pitch not a multiple of 16
    assert b'pitch' in lib.kt_last_error()
    assert lib.kt_spmv(None, None, None, None) < 0
    with pytest.raises(_lib.KtError):
        _lib.check(lib.kt_dot(None, None, None, None))


def test_preload_list_names_every_kernel(lib_path):
    """kt_preload_kernels walks csrc/kt_preload_list.inc; a kernel missing from it could dead-lock the in-process multi-rank
    mode in CUDA's lazy loader.  The list must be exactly the kernels of the built library (tools/gen_preload.py)."""
    elf = subprocess.run(['cuobjdump', '-elf', lib_path], capture_output=True, text=True).stdout
    mangled = sorted({l[len('.text.'):].strip() for l in elf.splitlines() if l.startswith('.text.')})
    dem = subprocess.run(['c++filt'], input='\n'.join(mangled), capture_output=True, text=True).stdout.splitlines()
    built = sorted({re.sub(r'^void ', '', d.strip()).split('(')[0] for d in dem})
    inc = open(os.path.join(ROOT, 'karman_taichi_b200', 'csrc', 'kt_preload_list.inc')).read()
    listed = sorted(re.findall(r'^KT_K\((.*)\)$', inc, flags=re.M))
    assert len(built) > 100 and listed == built


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'karman_taichi_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f
                assert 'from .. import oracle' not in txt


def test_no_gpu_means_loud_failure(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import karman_taichi_b200 as kt
    with pytest.raises(kt.KtError):
        kt.Field((8, 8))
