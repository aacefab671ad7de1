"""ctypes wrapper of oracle/specfab_oracle.c (the dense C restatement; TEST/BASELINE ONLY).

build():  gcc -O2 -std=c11 -fcx-fortran-rules -fopenmp -shared -fPIC  ->  oracle/_build/liboracle.so
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "specfab_oracle.c")
LIB = os.path.join(HERE, "_build", "liboracle.so")


class Opts(C.Structure):
    _fields_ = [("dt", C.c_double), ("iota", C.c_double), ("zeta", C.c_double), ("nu_mult", C.c_double),
                ("gamma0", C.c_double), ("lambda_", C.c_double),
                ("use_lrot", C.c_int), ("use_ddrx", C.c_int), ("use_cdrx", C.c_int), ("use_reg", C.c_int), ("rk4", C.c_int)]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fcx-fortran-rules", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None
_L = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_init.restype = C.c_int
        _lib.orc_init.argtypes = [C.c_int] + [C.c_void_p] * 4
        _lib.orc_step_batch.restype = C.c_int
        _lib.orc_step_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(Opts), C.c_int]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_set_num_threads.argtypes = [C.c_int]
        # use every host core we may run on (torchrun exports OMP_NUM_THREADS=1 for multi-rank launches)
        try:
            ncpu = len(os.sched_getaffinity(0))
        except AttributeError:
            ncpu = os.cpu_count() or 1
        _lib.orc_set_num_threads(ncpu)
    return _lib


def init(L):
    global _L
    import specfab_oracle as py
    T = py.load_tables()
    arrs = [np.ascontiguousarray(T[k]) for k in ("GC", "GCm", "GC_m1", "GC_p1")]
    n = lib().orc_init(L, *[a.ctypes.data for a in arrs])
    if n < 0:
        raise ValueError("bad L")
    _L = L
    return n


def num_threads():
    return lib().orc_num_threads()


def step_batch(nlm, ugrad, tau=None, dt=0.0, iota=1.0, zeta=0.0, nu_mult=1.0, Gamma0=0.0, Lambda=0.0,
               use_lrot=True, use_ddrx=False, use_cdrx=False, use_reg=True, scheme="euler", nsteps=1):
    """nlm (N,n) complex128, ugrad (N,3,3) [, tau (N,3,3)] -> new nlm (N,n).  OpenMP over nodes."""
    x = np.array(nlm, dtype=np.complex128, order="C")
    ug = np.ascontiguousarray(ugrad, dtype=np.float64)
    ta = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
    o = Opts(dt, iota, zeta, nu_mult, Gamma0, Lambda, int(use_lrot), int(use_ddrx), int(use_cdrx), int(use_reg), int(scheme == "rk4"))
    rc = lib().orc_step_batch(x.ctypes.data, x.shape[0], ug.ctypes.data, None if ta is None else ta.ctypes.data, C.byref(o), nsteps)
    if rc:
        raise RuntimeError("orc_step_batch failed")
    return x
