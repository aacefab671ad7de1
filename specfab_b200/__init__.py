"""specfab_b200 -- host-side mirror of the specfabpy interface for the fabric-evolution hot path.

Same names / argument meaning as the reference's f2py module (src/specfabpy.f90) for the path
SURVEY.md section 8 scopes, each with the batched `_arr` variant the reference's own convention
prescribes (leading node dimension, cf. Eij_tranisotropic_arr src/specfabpy.f90:474-486).
All arithmetic happens in libspecfab_b200.so (hand-written CUDA, sm_100a) through the C ABI of
include/specfab_b200.h; this file only marshals arrays.  No CPU fallback exists.

numpy (host) API: arrays shaped like the reference's, e.g. nlm (N, nlm_len) complex128,
ugrad (N,3,3); any memory order is accepted (converted to the node-contiguous Fortran order the
library wants, exactly what f2py does).  Device API (`*_dev`): torch CUDA tensors already in
library layout -- see `layout_nlm` / `layout_mat` -- run on torch's current stream.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (SFB_LROT, SFB_DDRX, SFB_CDRX, SFB_REG, SFB_EULER, SFB_RK4, SpecfabB200Error, StepOpts)

__all__ = ["init", "nlm_len", "step_arr", "step_arr_dev", "build_info", "layout_nlm", "layout_mat",
           "SFB_LROT", "SFB_DDRX", "SFB_CDRX", "SFB_REG", "SFB_EULER", "SFB_RK4", "SpecfabB200Error"]

_state = {"L": None, "n": None}


def init(L):
    """init(L) -> (lm[2,nlm_len], nlm_len)            reference: src/specfabpy.f90:150-160"""
    lib = _lib.load()
    _lib.check(lib.sfb_init(int(L)))
    n = lib.sfb_nlm_len()
    lm = np.zeros((n, 2), dtype=np.int32)
    _lib.check(lib.sfb_get_lm(lm.ctypes.data_as(C.c_void_p)))
    _state["L"], _state["n"] = int(L), n
    return lm.T.copy(), n


def nlm_len():
    """reference: src/specfabpy.f90:162-166"""
    return _lib.load().sfb_nlm_len()


def build_info():
    import json
    return json.loads(_lib.load().sfb_build_info().decode())


def _need_init():
    if _state["L"] is None:
        raise SpecfabB200Error(_lib.SFB_ENOINIT, "init(L) not called")
    return _state["n"]


def _terms(terms):
    if isinstance(terms, int):
        return terms
    m = {"lrot": SFB_LROT, "ddrx": SFB_DDRX, "cdrx": SFB_CDRX, "reg": SFB_REG}
    t = 0
    for s in terms:
        t |= m[s.lower()]
    return t


def _scheme(s):
    if s in (SFB_EULER, SFB_RK4):
        return s
    return {"euler": SFB_EULER, "rk4": SFB_RK4}[str(s).lower()]


def _opts(dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps, g0_ptr=None, lam_ptr=None):
    o = StepOpts()
    o.dt, o.iota, o.zeta, o.nu_mult = float(dt), float(iota), float(zeta), float(nu)
    o.gamma0 = 0.0 if g0_ptr else float(Gamma0)
    o.lambda_ = 0.0 if lam_ptr else float(Lambda)
    o.gamma0_arr, o.lambda_arr = g0_ptr, lam_ptr
    o.terms, o.scheme, o.nsteps = _terms(terms), _scheme(scheme), int(nsteps)
    return o


def _farr(a, dtype, shape_tail):
    a = np.asarray(a, dtype=dtype)
    if a.ndim != 1 + len(shape_tail) or tuple(a.shape[1:]) != tuple(shape_tail):
        raise ValueError("expected array of shape (N,%s), got %s" % (",".join(map(str, shape_tail)), a.shape))
    return np.asfortranarray(a)


def step_arr(nlm, ugrad, tau=None, dt=0.0, iota=1.0, zeta=0.0, nu=1.0, Gamma0=0.0, Lambda=0.0,
             terms=("lrot", "reg"), scheme="euler", nsteps=1):
    """Batched fused time step of N independent nodes (host arrays).

    nlm (N,nlm_len) complex128, ugrad (N,3,3), tau (N,3,3) or None (tau := sym(ugrad)).
    Gamma0 / Lambda: scalars or (N,) arrays.  Returns the new nlm (N,nlm_len), Fortran-ordered.
    Batches  nlm + dt*matmul(M_LROT + Gamma0*M_DDRX + Lambda*M_CDRX + M_REG, nlm)
    (reference per node: src/specfabpy/integrator.py:73-77, src/dynamics.f90:99-110)."""
    n = _need_init()
    lib = _lib.load()
    nlm_f = _farr(nlm, np.complex128, (n,))
    N = nlm_f.shape[0]
    ug = _farr(ugrad, np.float64, (3, 3))
    if ug.shape[0] != N:
        raise ValueError("ugrad has %d nodes, nlm has %d" % (ug.shape[0], N))
    ta = None
    if tau is not None:
        ta = _farr(tau, np.float64, (3, 3))
    keep = []

    def vec(x):
        if np.ndim(x) == 0:
            return None
        v = np.ascontiguousarray(x, dtype=np.float64)
        if v.shape != (N,):
            raise ValueError("per-node rate factor must have shape (N,)")
        keep.append(v)
        return v.ctypes.data

    o = _opts(dt, iota, zeta, nu, 0.0 if np.ndim(Gamma0) else Gamma0, 0.0 if np.ndim(Lambda) else Lambda,
              terms, scheme, nsteps, vec(Gamma0), vec(Lambda))
    out = np.empty((N, n), dtype=np.complex128, order="F")
    _lib.check(lib.sfb_step_arr(nlm_f.ctypes.data, out.ctypes.data, N, N, ug.ctypes.data,
                                ta.ctypes.data if ta is not None else None, C.byref(o)))
    return out


# ------------------------------------------------------------------------------------------
# device-resident API (torch tensors as memory handles; kernels run on torch's current stream)
# ------------------------------------------------------------------------------------------

def layout_nlm(nlm_t):
    """(N, nlm_len) complex tensor -> library layout (nlm_len, N) complex128, contiguous."""
    import torch
    return nlm_t.to(torch.complex128).t().contiguous()


def layout_mat(m_t):
    """(N,3,3) real tensor -> library layout (3,3,N): element [k,i,p] = m[p,i,k] (Fortran (N,3,3))."""
    import torch
    return m_t.to(torch.float64).permute(2, 1, 0).contiguous()


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def step_arr_dev(nlm, ugrad, tau=None, out=None, dt=0.0, iota=1.0, zeta=0.0, nu=1.0, Gamma0=0.0, Lambda=0.0,
                 terms=("lrot", "reg"), scheme="euler", nsteps=1):
    """Device-resident fused step.  nlm: (nlm_len, N) complex128 CUDA tensor (library layout),
    ugrad/tau: (3,3,N) float64 CUDA tensors (layout_mat).  out defaults to in-place.
    Gamma0/Lambda: scalars or (N,) float64 CUDA tensors.  Asynchronous on the current stream."""
    import torch
    n = _need_init()
    lib = _lib.load()
    if nlm.dtype != torch.complex128 or not nlm.is_cuda or not nlm.is_contiguous() or nlm.shape[0] != n:
        raise ValueError("nlm must be a contiguous CUDA complex128 tensor of shape (nlm_len, N)")
    N = nlm.shape[1]
    if out is None:
        out = nlm
    for t in (ugrad, tau):
        if t is not None and (t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != (3, 3, N)):
            raise ValueError("ugrad/tau must be contiguous CUDA float64 tensors of shape (3,3,N)")

    def vec(x):
        if not torch.is_tensor(x):
            return None
        if x.dtype != torch.float64 or not x.is_cuda or tuple(x.shape) != (N,):
            raise ValueError("per-node rate factor must be a CUDA float64 tensor of shape (N,)")
        return x.data_ptr()

    g0p, lamp = vec(Gamma0), vec(Lambda)
    o = _opts(dt, iota, zeta, nu, 0.0 if g0p else Gamma0, 0.0 if lamp else Lambda, terms, scheme, nsteps, g0p, lamp)
    _lib.check(lib.sfb_step_arr_dev(nlm.data_ptr(), out.data_ptr(), N, N, N, ugrad.data_ptr(), N,
                                    tau.data_ptr() if tau is not None else None, N, C.byref(o), _stream_ptr()))
    return out
