"""ctypes binding of libspecfab_b200.so (the C ABI of include/specfab_b200.h).

The library is the product; this module only loads it.  There is no Python/CPU fallback: if the
shared object is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspecfab_b200.so")

SFB_OK, SFB_EINVAL, SFB_ENOINIT, SFB_ECUDA, SFB_ENOTBUILT = 0, -1, -2, -3, -4
SFB_LROT, SFB_DDRX, SFB_CDRX, SFB_REG = 1, 2, 4, 8
SFB_EULER, SFB_RK4 = 1, 4
ST_TAYLOR_FALLBACK, ST_TAYLOR_FAILED, ST_NONFINITE = 1, 2, 4


class SpecfabB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("specfab_b200 error %d: %s" % (code, msg))
        self.code = code


class StepOpts(C.Structure):
    _fields_ = [("dt", C.c_double), ("iota", C.c_double), ("zeta", C.c_double), ("nu_mult", C.c_double),
                ("gamma0", C.c_double), ("lambda_", C.c_double),
                ("gamma0_arr", C.c_void_p), ("lambda_arr", C.c_void_p),
                ("terms", C.c_int32), ("scheme", C.c_int32), ("nsteps", C.c_int32), ("reserved", C.c_int32)]


_P = C.c_void_p
_I64 = C.c_int64
# name -> (restype, argtypes); must list every symbol declared in include/specfab_b200.h
SIGNATURES = {
    "sfb_init": (C.c_int, [C.c_int]),
    "sfb_nlm_len": (C.c_int, []),
    "sfb_get_lm": (C.c_int, [_P]),
    "sfb_finalize": (None, []),
    "sfb_last_error": (C.c_char_p, []),
    "sfb_build_info": (C.c_char_p, []),
    "sfb_step_arr": (C.c_int, [_P, _P, _I64, _I64, _P, _P, C.POINTER(StepOpts)]),
    "sfb_step_arr_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _I64, C.POINTER(StepOpts), _P]),
    "sfb_apply_bounds_rnlm_arr_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P]),
    "sfb_step_rnlm_arr": (C.c_int, [_P, _P, _I64, _I64, _P, _P, C.POINTER(StepOpts)]),
    "sfb_step_arr_multi": (C.c_int, [_P, _P, _I64, _I64, _P, _P, C.POINTER(StepOpts), C.POINTER(C.c_int), C.c_int]),
    "sfb_step_rnlm_arr_multi": (C.c_int, [_P, _P, _I64, _I64, _P, _P, C.POINTER(StepOpts), C.POINTER(C.c_int), C.c_int]),
    "sfb_step_rnlm_arr_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _I64, C.POINTER(StepOpts), _P]),
    "sfb_a2_arr": (C.c_int, [_P, _I64, _I64, _P]),
    "sfb_a2_arr_dev": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "sfb_a4_arr": (C.c_int, [_P, _I64, _I64, _P]),
    "sfb_a4_arr_dev": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "sfb_eig_arr": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "sfb_eig_arr_dev": (C.c_int, [_P, _I64, _I64, _P, _P, _P]),
    "sfb_eigframe_arr": (C.c_int, [_P, _I64, C.c_char_p, _P, _P]),
    "sfb_eigframe_arr_dev": (C.c_int, [_P, _I64, _I64, C.c_char_p, _P, _P, _P]),
    "sfb_Eij_tranisotropic_arr": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _P, C.c_double, C.c_int, _P, _P]),
    "sfb_Eij_tranisotropic_arr_dev": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _P, C.c_double, C.c_int, _P, _P, _P]),
    "sfb_Evw_tranisotropic_arr": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _P, C.c_double, C.c_int, _P, _P]),
    "sfb_Evw_tranisotropic_arr_dev": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _P, C.c_double, C.c_int, _P, _P, _P]),
    "sfb_a6_arr": (C.c_int, [_P, _I64, _I64, _P]),
    "sfb_a6_arr_dev": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "sfb_E_CAFFE_arr": (C.c_int, [_P, _I64, _I64, _P, C.c_double, C.c_double, C.c_int, _P]),
    "sfb_E_CAFFE_arr_dev": (C.c_int, [_P, _I64, _I64, _P, C.c_double, C.c_double, C.c_int, _P, _P]),
    "sfb_pfJ_arr": (C.c_int, [_P, _I64, _I64, C.c_int, _P]),
    "sfb_pfJ_arr_dev": (C.c_int, [_P, _I64, _I64, C.c_int, _P, _P]),
    "sfb_reduce_M_arr": (C.c_int, [_P, C.c_int, _I64, _P, _P, _P, _P]),
    "sfb_reduce_M_arr_dev": (C.c_int, [_P, C.c_int, _I64, _I64, _P, _P, _P, _P, _P]),
    "sfb_M_LROT_reduced_arr": (C.c_int, [_P, _P, _I64, C.c_double, C.c_double, _P, _P, _P, _P]),
    "sfb_M_LROT_reduced_arr_dev": (C.c_int, [_P, _P, _I64, _I64, C.c_double, C.c_double, _P, _P, _P, _P, _P]),
    "sfb_M_DDRX_reduced_arr": (C.c_int, [_P, _I64, _P, _I64, C.c_int, _P, _P, _P, _P]),
    "sfb_M_DDRX_reduced_arr_dev": (C.c_int, [_P, _I64, _P, _I64, _I64, C.c_int, _P, _P, _P, _P, _P]),
    "sfb_ai_to_nlm_arr": (C.c_int, [C.c_int, _P, _I64, _P]),
    "sfb_ai_to_nlm_arr_dev": (C.c_int, [C.c_int, _P, _I64, _I64, _P, _I64, _P]),
    "sfb_step_moments_Eij_arr_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _I64, _P, _P, C.c_double, C.c_int,
                                               _P, _P, _P, _P, _P, _P, _P]),
    "sfb_Eij_orthotropic_arr": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P, _P, _P, C.c_double, C.c_int, _P]),
    "sfb_Eij_orthotropic_arr_dev": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I64, _P, _P, _P, _P, C.c_double, C.c_int, _P, _P]),
    "sfb_Eij_eigenframe_arr": (C.c_int, [_P, _I64, _I64, _P, C.c_double, C.c_int, _P, _P, _P, _P]),
    "sfb_Eij_eigenframe_arr_dev": (C.c_int, [_P, _I64, _I64, _P, C.c_double, C.c_int, _P, _P, _P, _P, _P]),
    "sfb_Eij_eigenframe_rnlm_arr_dev": (C.c_int, [_P, _I64, _I64, _P, C.c_double, C.c_int, _P, _P, _P, _P, _P, _P]),
    "sfb_step_moments_Eij_rnlm_arr_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _I64, _P, _P, C.c_double, C.c_int,
                                                    _P, _P, _P, _P, _P, _P]),
    "sfb_M_LROT_arr": (C.c_int, [_P, _P, _I64, C.c_double, C.c_double, _P]),
    "sfb_M_LROT_arr_dev": (C.c_int, [_P, _P, _I64, _I64, C.c_double, C.c_double, _P, _P]),
    "sfb_M_DDRX_src_arr": (C.c_int, [_P, _I64, _P]),
    "sfb_M_DDRX_src_arr_dev": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "sfb_M_DDRX_arr": (C.c_int, [_P, _I64, _P, _I64, _P]),
    "sfb_M_DDRX_arr_dev": (C.c_int, [_P, _I64, _P, _I64, _I64, _P, _P]),
    "sfb_M_REG_arr": (C.c_int, [_P, _I64, _P]),
    "sfb_M_REG_arr_dev": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "sfb_M_CDRX": (C.c_int, [_P]),
    "sfb_apply_bounds_arr": (C.c_int, [_P, _P, _I64, _I64]),
    "sfb_apply_bounds_arr_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P]),
    "sfb_rnlm_len": (C.c_int, []),
    "sfb_nlm_to_rnlm_arr": (C.c_int, [_P, _P, _I64]),
    "sfb_nlm_to_rnlm_arr_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P]),
    "sfb_rnlm_to_nlm_arr": (C.c_int, [_P, _P, _I64]),
    "sfb_rnlm_to_nlm_arr_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P]),
    "sfb_set_variant": (C.c_int, [C.c_int]),
    "sfb_dev_malloc": (C.c_int, [C.POINTER(_P), _I64]),
    "sfb_dev_free": (C.c_int, [_P]),
    "sfb_memcpy_h2d": (C.c_int, [_P, _P, _I64]),
    "sfb_memcpy_d2h": (C.c_int, [_P, _P, _I64]),
    "sfb_host_alloc_pinned": (C.c_int, [C.POINTER(_P), _I64]),
    "sfb_host_free_pinned": (C.c_int, [_P]),
    "sfb_host_register": (C.c_int, [_P, _I64]),
    "sfb_host_unregister": (C.c_int, [_P]),
    "sfb_sync": (C.c_int, []),
    "sfb_device_count": (C.c_int, []),
    "sfb_set_device": (C.c_int, [C.c_int]),
}

_lib = None


def load():
    """Load the shared library (raises if it has not been built: python -m specfab_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpecfabB200Error(SFB_ENOTBUILT, "%s not found -- build it with `python -m specfab_b200.build` "
                                   "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != SFB_OK:
        raise SpecfabB200Error(rc, (load().sfb_last_error() or b"").decode())
    return rc
