"""ctypes binding of libkarman_b200.so (the C ABI in ``include/karman_b200.h``).

There is no CPU fallback: if the shared library is missing or a CUDA device is not usable, the
calls raise.  The library is built in-tree by ``karman_taichi_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# KT_LIB_PATH: development override to load a tuning variant of the same library (tools/variants.py)
LIB_PATH = os.environ.get('KT_LIB_PATH') or os.path.join(HERE, 'libkarman_b200.so')

KT_F64, KT_F32 = 0, 1

STATUS_CONVERGED = 0
STATUS_MAX_ITER = 1
STATUS_RHO_ZERO = 2
STATUS_ALPHA_DIV_ZERO = 3
STATUS_OMEGA_DIV_ZERO = 4
STATUS_NON_FINITE = 5
STATUS_OMEGA_ZERO = 6
STATUS_COMM_TIMEOUT = 7

FLAG_WARM_START = 1
FLAG_OMEGA_ZERO_BREAK = 2


class KtError(RuntimeError):
    pass


class KrylovWs(C.Structure):
    """``kt_krylov_ws``"""
    _fields_ = [('r', C.c_void_p), ('p', C.c_void_p), ('Ap', C.c_void_p), ('r_tld', C.c_void_p),
                ('s', C.c_void_p), ('t', C.c_void_p), ('history', C.c_void_p), ('history_cap', C.c_long),
                ('p2', C.c_void_p), ('Ap2', C.c_void_p)]


class P2PComm(C.Structure):
    """``kt_p2p_comm``"""
    _fields_ = [('rank', C.c_int), ('nranks', C.c_int), ('peer_halo_recv', C.c_void_p * 2), ('halo_recv', C.c_void_p),
                ('comm_block', C.c_void_p * 16), ('local_dots', C.c_void_p)]


class Geom(C.Structure):
    """``kt_geom``"""
    _fields_ = [('NX', C.c_int), ('NY', C.c_int), ('pitch', C.c_int), ('row_lo', C.c_int), ('row_hi', C.c_int),
                ('dtype', C.c_int), ('num_sms', C.c_int)]


class FrameParams(C.Structure):
    """``kt_frame_params``"""
    _fields_ = [('nx', C.c_int), ('ny', C.c_int), ('pitch_u', C.c_int), ('pitch_v', C.c_int), ('pitch_p', C.c_int),
                ('scale', C.c_int), ('monolith', C.c_int), ('dx', C.c_double), ('dy', C.c_double)]


class ChannelParams(C.Structure):
    """``kt_channel_params``"""
    _fields_ = [('nx', C.c_int), ('ny', C.c_int),
                ('dx', C.c_double), ('dy', C.c_double), ('dt', C.c_double), ('rho', C.c_double), ('mu', C.c_double),
                ('velo_rel', C.c_double), ('p_rel', C.c_double),
                ('cyl_cx', C.c_int), ('cyl_cy', C.c_int), ('cyl_r2', C.c_int),
                ('pitch_u', C.c_int), ('pitch_v', C.c_int), ('pitch_p', C.c_int),
                ('row0', C.c_int), ('row_lo', C.c_int), ('row_hi', C.c_int), ('rows_u', C.c_int), ('rows_p', C.c_int)]


class SimpleParams(C.Structure):
    """``kt_simple_params``"""
    _fields_ = [('nx', C.c_int), ('ny', C.c_int),
                ('dx', C.c_double), ('dy', C.c_double), ('dt', C.c_double), ('rho', C.c_double), ('mu', C.c_double),
                ('alpha_p', C.c_double), ('alpha_u', C.c_double), ('alpha_m', C.c_double),
                ('bc_w', C.c_double * 2), ('bc_e', C.c_double * 2), ('bc_n', C.c_double * 2), ('bc_s', C.c_double * 2),
                ('pin_p', C.c_int), ('east_wall_bug', C.c_int),
                ('pitch_u', C.c_int), ('pitch_v', C.c_int), ('pitch_p', C.c_int)]


_P, _I, _D, _L = C.c_void_p, C.c_int, C.c_double, C.c_long
_PD, _PL, _PI, _PF = C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_int), C.POINTER(C.c_float)
_WS, _SP, _CP = C.POINTER(KrylovWs), C.POINTER(SimpleParams), C.POINTER(ChannelParams)

# name -> (restype, argtypes); every symbol include/karman_b200.h declares
SIGNATURES = {
    'kt_version': (_I, []),
    'kt_last_error': (C.c_char_p, []),
    'kt_pitch_for': (_I, [_I, _I]),
    'kt_create': (_I, [_I, _I, _I, _I, _I, _P, C.POINTER(_P)]),
    'kt_destroy': (_I, [_P]),
    'kt_set_stream': (_I, [_P, _P]),
    'kt_sync': (_I, [_P]),
    'kt_set_rows': (_I, [_P, _I, _I]),
    'kt_launch_count': (C.c_longlong, [_P]),
    'kt_set_cluster': (_I, [_P, _I]),
    'kt_preload_kernels': (_I, [_P]),
    'kt_upload_field': (_I, [_P, _P, _P]),
    'kt_download_field': (_I, [_P, _P, _P]),
    'kt_upload_coef': (_I, [_P, _P, _P, _P, C.c_size_t]),
    'kt_download_coef': (_I, [_P, _P, _P, _P, C.c_size_t]),
    'kt_coef_aos_to_soa': (_I, [_P, _P, _P]),
    'kt_coef_soa_to_aos': (_I, [_P, _P, _P]),
    'kt_spmv': (_I, [_P, _P, _P, _P]),
    'kt_spmv_dot': (_I, [_P, _P, _P, _P, _P, _PD]),
    'kt_bicg_s_t': (_I, [_P, _P, _P, _P, _D, _P, _P, _PD]),
    'kt_dot': (_I, [_P, _P, _P, _PD]),
    'kt_copy': (_I, [_P, _P, _P]),
    'kt_fill': (_I, [_P, _P, _D]),
    'kt_axpy': (_I, [_P, _P, _D, _P]),
    'kt_xpay': (_I, [_P, _P, _P, _D, _P]),
    'kt_xmay': (_I, [_P, _P, _P, _D, _P]),
    'kt_bicg_update_p': (_I, [_P, _P, _P, _P, _D, _D]),
    'kt_bicg_update_x': (_I, [_P, _P, _P, _P, _D, _D]),
    'kt_bicg_xr': (_I, [_P, _P, _P, _P, _P, _P, _P, _D, _D, _PD]),
    'kt_cg_xr': (_I, [_P, _P, _P, _P, _P, _D, _PD]),
    'kt_cg_solve': (_I, [_P, _P, _P, _P, _WS, _D, _L, _PL, _PD, _PI]),
    'kt_bicgstab_solve': (_I, [_P, _P, _P, _P, _P, _WS, _D, _L, _I, _PL, _PD, _PI]),
    'kt_bicgstab_bench': (_I, [_P, _P, _P, _P, _WS, _L, _PF]),
    'kt_cg_bench': (_I, [_P, _P, _P, _P, _WS, _L, _PF]),
    'kt_bicgstab_solve_host': (_I, [_P, _P, _P, _P, _P, _P, C.c_size_t, _P, _P, _WS, _D, _L, _PL, _PD, _PI]),
    'kt_dist_begin': (_I, [_P, _D, _L, _I]),
    'kt_dist_bicg_phase': (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'kt_dist_scalar': (_I, [_P, _I, _P, _P, _L]),
    'kt_dist_poll': (_I, [_P, _PI, _PL, _PI, _PD]),
    'kt_p2p_comm_block_bytes': (_I, []),
    'kt_ipc_alloc': (_I, [_P, C.c_size_t, C.POINTER(_P), C.c_char_p]),
    'kt_ipc_open': (_I, [_P, C.c_char_p, C.POINTER(_P)]),
    'kt_ipc_close': (_I, [_P, _P]),
    'kt_ipc_free': (_I, [_P, _P]),
    'kt_enable_peer_access': (_I, [_P, _I]),
    'kt_p2p_bicgstab_solve': (_I, [_P, _P, _P, _P, _P, _WS, C.POINTER(P2PComm), _D, _L, _I, _I, _PL, _PD, _PI]),
    'kt_p2p_halo_slots': (_I, []),
    'kt_p2p_exchange': (_I, [_P, C.POINTER(P2PComm), _P]),
    'kt_p2p_allreduce': (_I, [_P, C.POINTER(P2PComm), _P, _PD, _I, _PD]),
    'kt_p2p_cg_solve': (_I, [_P, _P, _P, _P, _WS, C.POINTER(P2PComm), _D, _L, _I, _PL, _PD, _PI]),
    'kt_reduce_plan': (_I, [C.POINTER(Geom), _I, _I, C.POINTER(C.c_longlong)]),
    'kt_resident_plan': (_I, [C.POINTER(Geom), _I, _I, _I, _I, C.POINTER(C.c_longlong)]),
    'kt_get_geom': (_I, [_P, C.POINTER(Geom)]),
    'kt_prof_begin': (_I, [_P, _I]),
    'kt_prof_end': (_I, [_P, _PD, _PL, _I]),
    'kt_assemble_uv': (_I, [_P, _SP, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'kt_set_bc_cavity': (_I, [_P, _SP, _P, _P, _P, _P, _P, _P]),
    'kt_assemble_p': (_I, [_P, _SP, _P, _P, _P, _P, _P, _P, _P]),
    'kt_relax_uv': (_I, [_P, _SP, _P, _P, _P, _P, _PD]),
    'kt_correct_puv': (_I, [_P, _SP, _P, _P, _P, _P, _P, _P]),
    'kt_mdiv_max': (_I, [_P, _SP, _P, _P, _P, _PD]),
    'kt_channel_init': (_I, [_P, _CP, _P, _P, _P, _P, _P, _P]),
    'kt_channel_assemble_uv': (_I, [_P, _CP] + [_P] * 14),
    'kt_channel_assemble_p': (_I, [_P, _CP] + [_P] * 9),
    'kt_channel_relax': (_I, [_P, _CP, _P, _P, _P, _P]),
    'kt_outlet_scale': (_I, [_P, _CP, _P, _P, _P, _PD]),
    'kt_outlet_sums': (_I, [_P, _CP, _P, _P]),
    'kt_outlet_apply': (_I, [_P, _CP, _P, _D, _D]),
    'kt_channel_udiv': (_I, [_P, _CP, _P, _P, _P]),
    'kt_channel_correct': (_I, [_P, _CP, _P, _P, _P, _P, _P, _P, _P, _PD]),
    'kt_display_frame': (_I, [_P, C.POINTER(FrameParams), _P, _P, _P, _P, _P, _P, _PD]),
    'kt_jacobi_sweep': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _D, _PD]),
}

_lib = None


def load():
    """dlopen the library and declare every prototype.  Raises ``KtError`` if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KtError(f'{LIB_PATH} is missing: build it with `python -m karman_taichi_b200.build` '
                      '(there is no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().kt_last_error().decode('utf-8', 'replace')
        raise KtError(f'libkarman_b200 call failed (rc={rc}): {msg}')


def pitch_for(ny, dtype_code=KT_F64):
    """``kt_pitch_for``: the row pitch (elements) the kernels run fastest with -- rows padded to 512 bytes."""
    return int(load().kt_pitch_for(int(ny), int(dtype_code)))
