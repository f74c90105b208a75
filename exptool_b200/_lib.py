"""
_lib.py -- ctypes binding of libbfe.so (the C ABI declared in include/bfe.h).

There is no CPU fallback: importing this module without the built library
raises, and every op raises if CUDA is unavailable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BFE_LIB') or os.path.join(_HERE, 'libbfe.so')      # BFE_LIB: A/B builds (profiles/)


class BfeError(RuntimeError):
    pass


class EofParams(C.Structure):
    _fields_ = [('mmax', C.c_int32), ('norder', C.c_int32), ('numx', C.c_int32), ('numy', C.c_int32),
                ('cmap', C.c_int32), ('dens', C.c_int32),
                ('xmin', C.c_double), ('dx', C.c_double), ('ymin', C.c_double), ('dy', C.c_double),
                ('ascale', C.c_double), ('hscale', C.c_double)]


class SlParams(C.Structure):
    _fields_ = [('lmax', C.c_int32), ('nmax', C.c_int32), ('numr', C.c_int32), ('cmap', C.c_int32),
                ('scale', C.c_double)]


_P = C.c_void_p       # device pointers and streams travel as plain addresses
_I64 = C.c_int64
_INT = C.c_int
_DBL = C.c_double

# name -> (restype, argtypes); must list every symbol include/bfe.h declares
SIGNATURES = {
    'bfe_error_string': (C.c_char_p, [_INT]),
    'bfe_last_cuda_error': (C.c_char_p, []),
    'bfe_version': (_INT, []),
    'bfe_launch_count': (C.c_uint64, []),
    'bfe_set_option': (_INT, [C.c_char_p, _INT]),
    'bfe_get_option': (_INT, [C.c_char_p]),
    'bfe_kernel_time_ms': (C.c_double, [C.c_char_p]),
    'bfe_fp64_peak': (_INT, [_INT, C.POINTER(C.c_double), _P]),
    'bfe_eof_set_table_fp32': (_INT, [_P, _INT]),
    'bfe_sl_set_table_fp32': (_INT, [_P, _INT]),
    'bfe_eof_create': (_INT, [C.POINTER(EofParams)] + [_P] * 6 + [_P, C.POINTER(_P)]),
    'bfe_eof_destroy': (None, [_P]),
    'bfe_eof_clone': (_INT, [_P, _P, C.POINTER(_P)]),
    'bfe_eof_accumulate': (_INT, [_P, _I64] + [_P] * 4 + [_P, _P, _P]),
    'bfe_eof_prepare': (_INT, [_P, _I64] + [_P] * 4 + [_P]),
    'bfe_eof_accumulate_prepared': (_INT, [_P, _P, _P, _P]),
    'bfe_eof_force_prepared': (_INT, [_P] + [_P] * 6 + [_P]),
    'bfe_eof_accumulate_host': (_INT, [_P, _I64] + [_P] * 4 + [_P, _P, _P]),
    'bfe_eof_force_host': (_INT, [_P, _I64] + [_P] * 3 + [_P] * 6 + [_P]),
    'bfe_eof_contract': (_INT, [_P, _P, _P, _INT, _INT, _INT, _INT, _P]),
    'bfe_eof_force_contracted': (_INT, [_P, _I64] + [_P] * 3 + [_P] * 6 + [_P]),
    'bfe_eof_force': (_INT, [_P, _I64] + [_P] * 3 + [_P, _P, _INT, _INT, _INT, _INT] + [_P] * 6 + [_P]),
    'bfe_eof_force_eval_points': (_INT, [_P, _I64] + [_P] * 3 + [_P] * 5 + [_P]),
    'bfe_sl_create': (_INT, [C.POINTER(SlParams)] + [_P] * 5 + [_P, C.POINTER(_P)]),
    'bfe_sl_destroy': (None, [_P]),
    'bfe_sl_accumulate': (_INT, [_P, _I64] + [_P] * 4 + [_INT, _P, _P]),
    'bfe_sl_contract': (_INT, [_P, _P, _INT, _INT, _INT, _INT, _P]),
    'bfe_sl_accumulate_host': (_INT, [_P, _I64] + [_P] * 4 + [_INT, _P, _P]),
    'bfe_sl_force_host': (_INT, [_P, _I64] + [_P] * 3 + [_P] * 6 + [_P]),
    'bfe_sl_force_contracted': (_INT, [_P, _I64] + [_P] * 3 + [_P] * 6 + [_P]),
    'bfe_sl_force': (_INT, [_P, _I64] + [_P] * 3 + [_P, _INT, _INT, _INT] + [_P] * 6 + [_P]),
    'bfe_peer_buffer_create': (_INT, [_I64, C.POINTER(_P), C.c_char_p]),
    'bfe_peer_buffer_open': (_INT, [C.c_char_p, C.POINTER(_P)]),
    'bfe_peer_buffer_close': (_INT, [_P]),
    'bfe_peer_buffer_destroy': (_INT, [_P]),
    'bfe_peer_create': (_INT, [_INT, _INT, _I64, C.POINTER(_P), C.POINTER(_P)]),
    'bfe_peer_destroy': (None, [_P]),
    'bfe_peer_allreduce': (_INT, [_P, _P, _I64, _P]),
    'bfe_peer_error': (_INT, [_P, _P, C.POINTER(C.c_uint64)]),
    'bfe_peer_poison': (_INT, [_P, C.c_uint64, _P]),
    'bfe_eof_return_bins': (_INT, [C.POINTER(EofParams), _I64] + [_P] * 6 + [_P]),
    'bfe_eof_get_pot': (_INT, [_P, _I64, _P, _P, C.c_double, _P, _P, _P]),
    'bfe_sl_radial_matrices': (_INT, [_P, _I64, _P, _P, _P, _P, _P]),
    'bfe_legendre_tables': (_INT, [_INT, _I64, _P, _P, _P, _P]),
    'bfe_sl_contract_density': (_INT, [_P, _P, _INT, _INT, _INT, _INT, _P]),
    'bfe_sl_density_contracted': (_INT, [_P, _I64] + [_P] * 3 + [_P] * 2 + [_P]),
    'bfe_sl_density_eval_points': (_INT, [_P, _I64] + [_P] * 3 + [_P] * 2 + [_P]),
    'bfe_sl_force_eval_points': (_INT, [_P, _I64] + [_P] * 3 + [_INT] + [_P] * 5 + [_P]),
    'bfe_field_force_cart': (_INT, [_P, _P, _I64] + [_P] * 3 + [_DBL, _P, _P]),
    'bfe_field_force_cyl': (_INT, [_P, _P, _I64] + [_P] * 3 + [_DBL, _P, _P]),
    'bfe_leapfrog': (_INT, [_P, _P, _I64, _I64, _DBL, _DBL, _P, _P, _I64, _INT, _INT, _P, _P]),
    'bfe_bar_fourier': (_INT, [_I64, _P, _P, _DBL, _DBL, _P, _P]),
    'bfe_affine_xy': (_INT, [_I64, _DBL, _DBL, _DBL, _DBL] + [_P] * 6 + [_P]),
    'bfe_inner_com': (_INT, [_I64] + [_P] * 7 + [_I64, _P, _P]),
    'bfe_leapfrog_dt': (_INT, [_P, _P, _I64, _I64, _P, _DBL, _P, _P, _I64, _INT, _INT, _P, _P]),
}

_lib = None


def load():
    """Load libbfe.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('exptool_b200: %s is missing -- build it with '
                          '`python exptool_b200/csrc/build.py` (there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        lib = load()
        msg = lib.bfe_error_string(rc).decode()
        if rc == -2:
            msg += ': ' + lib.bfe_last_cuda_error().decode()
        raise BfeError('libbfe: %s (code %d)' % (msg, rc))
