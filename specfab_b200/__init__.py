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

__all__ = ["apply_bounds", "apply_bounds_arr", "nlm_to_rnlm", "rnlm_to_nlm", "nlm_to_rnlm_arr", "rnlm_to_nlm_arr", "rnlm_len", "M_LROT", "M_DDRX", "M_DDRX_src", "M_CDRX", "M_REG", "dndt_LATROT", "dndt_DDRX", "dndt_CDRX", "dndt_REG", "reduce_M", "reduce_M_arr", "M_LROT_reduced_arr", "M_DDRX_reduced_arr", "M_LROT_arr", "M_DDRX_arr", "M_DDRX_src_arr", "M_REG_arr",
           "nlm_LROT", "init", "nlm_len", "step_arr", "step_arr_dev", "step_rnlm_arr", "step_rnlm_arr_dev", "step_moments_Eij_rnlm_arr_dev", "apply_bounds_rnlm_arr_dev", "pin_array", "unpin_array", "build_info", "layout_nlm", "layout_mat",
           "a2", "a4", "eig", "a2_arr", "a4_arr", "eig_arr", "eigframe_arr", "Eij_tranisotropic", "Eij_tranisotropic_arr",
           "a6", "a6_arr", "a2_to_nlm", "a4_to_nlm", "a6_to_nlm", "a2_to_nlm_arr", "a4_to_nlm_arr", "a6_to_nlm_arr", "E_CAFFE", "E_CAFFE_arr", "pfJ", "pfJ_arr", "Eij_eigenframe_arr", "Eij_orthotropic", "Eij_orthotropic_arr", "Eij_orthotropic_arr_dev", "a2_arr_dev", "Eij_eigenframe_arr_dev", "step_moments_Eij_arr_dev", "Eij_tranisotropic_arr_dev",
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


SFB_STEP_GENERAL = 1       # sfb_step_opts.reserved flag (include/specfab_b200.h)


def _opts(dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps, g0_ptr=None, lam_ptr=None, general=False):
    o = StepOpts()
    o.reserved = SFB_STEP_GENERAL if general else 0
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
             terms=("lrot", "reg"), scheme="euler", nsteps=1, out=None, devices=None, general=False):
    """Batched fused time step of N independent nodes (host arrays).

    devices: list of CUDA device ordinals -> the batch is sharded over these GPUs by contiguous node range
    (sfb_step_arr_multi: one host thread and staging ring per device); None = the current device.
    general=True: treat every state as a general complex vector (SFB_STEP_GENERAL: no real-ODF shortcut, results
    independent of how nodes are batched).

    nlm (N,nlm_len) complex128, ugrad (N,3,3), tau (N,3,3) or None (tau := sym(ugrad)).
    Gamma0 / Lambda: scalars or (N,) arrays.  Returns the new nlm (N,nlm_len), Fortran-ordered
    (written into `out` when given: a Fortran-ordered (N,nlm_len) complex128 array, e.g. pinned memory).
    Batches  nlm + dt*matmul(M_LROT + Gamma0*M_DDRX + Lambda*M_CDRX + M_REG, nlm)
    (reference per node: src/specfabpy/integrator.py:73-77, src/dynamics.f90:99-110)."""
    return _step_host(False, nlm, ugrad, tau, dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps, out, devices, general)


def step_rnlm_arr(rnlm, ugrad, tau=None, dt=0.0, iota=1.0, zeta=0.0, nu=1.0, Gamma0=0.0, Lambda=0.0,
                  terms=("lrot", "reg"), scheme="euler", nsteps=1, out=None, devices=None):
    """step_arr on REDUCED-FORM states: rnlm (N, rnlm_len) complex128 holds the m >= 0 coefficients of a real-valued ODF
    (nlm_to_rnlm / rnlm_to_nlm, src/reducedform.f90:160-187 -- the state representation of the FE couplers,
    src/specfabpy/fenics/CPO.py).  Same arguments and result as step_arr with every state array in reduced form;
    equals nlm_to_rnlm_arr(step_arr(rnlm_to_nlm_arr(rnlm), ...)) bit for bit and moves 25/45 (L=8) of the state bytes."""
    return _step_host(True, rnlm, ugrad, tau, dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps, out, devices, False)


def _step_host(reduced, nlm, ugrad, tau, dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps, out, devices=None, general=False):
    n = _need_init()
    lib = _lib.load()
    if reduced:
        n = rnlm_len()
    nlm_f = _farr(nlm, np.complex128, (n,))
    N = nlm_f.shape[0]
    ug = _farr(ugrad, np.float64, (3, 3))
    if ug.shape[0] != N:
        raise ValueError("ugrad has %d nodes, nlm has %d" % (ug.shape[0], N))
    ta = None
    if tau is not None:
        ta = _farr(tau, np.float64, (3, 3))
        if ta.shape[0] != N:
            raise ValueError("tau has %d nodes, nlm has %d" % (ta.shape[0], N))
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
              terms, scheme, nsteps, vec(Gamma0), vec(Lambda), general)
    if out is None:
        out = np.empty((N, n), dtype=np.complex128, order="F")
    elif out.shape != (N, n) or out.dtype != np.complex128 or not out.flags.f_contiguous:
        raise ValueError("out must be a Fortran-ordered complex128 array of shape (N, %s)" % ("rnlm_len" if reduced else "nlm_len"))
    if devices is not None:
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        _lib.check((lib.sfb_step_rnlm_arr_multi if reduced else lib.sfb_step_arr_multi)(
            nlm_f.ctypes.data, out.ctypes.data, N, N, ug.ctypes.data, ta.ctypes.data if ta is not None else None, C.byref(o), devs, len(devices)))
        return out
    _lib.check((lib.sfb_step_rnlm_arr if reduced else lib.sfb_step_arr)(nlm_f.ctypes.data, out.ctypes.data, N, N, ug.ctypes.data,
                                ta.ctypes.data if ta is not None else None, C.byref(o)))
    return out


# ------------------------------------------------------------------------------------------
# operators (matrix form), reference: src/specfabpy.f90:172-245
# ------------------------------------------------------------------------------------------

def M_LROT_arr(eps, omg, iota, zeta):
    """M_LROT of every node: eps, omg (N,3,3) -> (N,nlm_len,nlm_len) complex128 (Fortran order)"""
    n = _need_init()
    e, w = _farr(eps, np.float64, (3, 3)), _farr(omg, np.float64, (3, 3))
    N = e.shape[0]
    M = np.empty((N, n, n), dtype=np.complex128, order="F")
    _lib.check(_lib.load().sfb_M_LROT_arr(e.ctypes.data, w.ctypes.data, N, float(iota), float(zeta), M.ctypes.data))
    return M


def M_DDRX_src_arr(tau):
    n = _need_init()
    t = _farr(tau, np.float64, (3, 3))
    N = t.shape[0]
    M = np.empty((N, n, n), dtype=np.complex128, order="F")
    _lib.check(_lib.load().sfb_M_DDRX_src_arr(t.ctypes.data, N, M.ctypes.data))
    return M


def M_DDRX_arr(nlm, tau):
    n = _need_init()
    x = np.asfortranarray(np.asarray(nlm, dtype=np.complex128))
    t = _farr(tau, np.float64, (3, 3))
    N = t.shape[0]
    if x.shape != (N, n):
        raise ValueError("nlm must have shape (N, nlm_len)")
    M = np.empty((N, n, n), dtype=np.complex128, order="F")
    _lib.check(_lib.load().sfb_M_DDRX_arr(x.ctypes.data, N, t.ctypes.data, N, M.ctypes.data))
    return M


def M_REG_arr(eps):
    n = _need_init()
    e = _farr(eps, np.float64, (3, 3))
    N = e.shape[0]
    M = np.empty((N, n, n), dtype=np.float64, order="F")
    _lib.check(_lib.load().sfb_M_REG_arr(e.ctypes.data, N, M.ctypes.data))
    return M


def _four(N):
    r = rnlm_len()
    return [np.empty((N, r, r), dtype=np.float64, order="F") for _ in range(4)]


def reduce_M_arr(M):
    """reduce_M of every node: dense M (N,nlm_len,nlm_len) complex128 or float64 -> (Mrr, Mri, Mir, Mii), each
    (N,rnlm_len,rnlm_len)            reference per node: src/specfabpy.f90:1066-1074, src/reducedform.f90:76-120"""
    n = _need_init()
    M = np.asarray(M)
    cplx = np.iscomplexobj(M)
    m = _farr(M, np.complex128 if cplx else np.float64, (n, n))
    N = m.shape[0]
    o = _four(N)
    _lib.check(_lib.load().sfb_reduce_M_arr(m.ctypes.data, int(cplx), N, *[x.ctypes.data for x in o]))
    return tuple(o)


def reduce_M(M, rnlm_len_=None):
    """reference: src/specfabpy.f90:1066-1074 -> (Mrr, Mri, Mir, Mii)"""
    return tuple(np.ascontiguousarray(x[0]) for x in reduce_M_arr(np.asarray(M)[None]))


def M_LROT_reduced_arr(eps, omg, iota, zeta):
    """reduce_M(M_LROT) of every node without materialising the dense operator -> (Mrr, Mri, Mir, Mii)
    (what src/specfabpy/fenics/CPO.py:200-211 builds per node)"""
    _need_init()
    e, w = _farr(eps, np.float64, (3, 3)), _farr(omg, np.float64, (3, 3))
    N = e.shape[0]
    o = _four(N)
    _lib.check(_lib.load().sfb_M_LROT_reduced_arr(e.ctypes.data, w.ctypes.data, N, float(iota), float(zeta), *[x.ctypes.data for x in o]))
    return tuple(o)


def M_DDRX_reduced_arr(nlm, tau, src_only=False):
    """reduce_M(M_DDRX(nlm, tau)) (or of M_DDRX_src(tau) with src_only=True) of every node -> (Mrr, Mri, Mir, Mii)"""
    n = _need_init()
    t = _farr(tau, np.float64, (3, 3))
    N = t.shape[0]
    x = None
    if not src_only:
        x = np.asfortranarray(np.asarray(nlm, dtype=np.complex128))
        if x.shape != (N, n):
            raise ValueError("nlm must have shape (N, nlm_len)")
    o = _four(N)
    _lib.check(_lib.load().sfb_M_DDRX_reduced_arr(x.ctypes.data if x is not None else None, N, t.ctypes.data, N, int(bool(src_only)),
                                                  *[y.ctypes.data for y in o]))
    return tuple(o)


def M_LROT(nlm, eps, omg, iota, zeta):
    """reference: src/specfabpy.f90:172-180 (nlm only fixes the size)"""
    return np.ascontiguousarray(M_LROT_arr(np.asarray(eps)[None], np.asarray(omg)[None], iota, zeta)[0])


def M_DDRX(nlm, tau):
    """reference: src/specfabpy.f90:182-190 (multiply by Gamma0 yourself, like the reference)"""
    return np.ascontiguousarray(M_DDRX_arr(np.asarray(nlm)[None, :], np.asarray(tau)[None])[0])


def M_DDRX_src(nlm, tau):
    """reference: src/specfabpy.f90:192-199"""
    return np.ascontiguousarray(M_DDRX_src_arr(np.asarray(tau)[None])[0])


def M_CDRX(nlm):
    """reference: src/specfabpy.f90:228-235 (real diagonal; returned complex like the f2py wrapper)"""
    n = _need_init()
    M = np.zeros((n, n), dtype=np.float64)
    _lib.check(_lib.load().sfb_M_CDRX(M.ctypes.data))
    return M.astype(np.complex128)


def M_REG(nlm, eps):
    """reference: src/specfabpy.f90:237-245"""
    return np.ascontiguousarray(M_REG_arr(np.asarray(eps)[None])[0])


# names of older specfab releases for the same operators (BASELINE.json north_star; SURVEY.md naming note)
dndt_LATROT = M_LROT
dndt_DDRX = M_DDRX
dndt_CDRX = M_CDRX
dndt_REG = M_REG


def nlm_LROT(nlm0, dt, Nt, D, W, iota):
    """Euler integrator of lattice rotation for one parcel with time-dependent D(t), W(t) -> nlm (Nt, nlm_len)
    reference: src/specfabpy.f90:260-268, src/dynamics.f90:99-110 (zeta = 0, no regularisation)"""
    n = _need_init()
    out = np.zeros((Nt, n), dtype=np.complex128)
    out[0] = nlm0
    cur = np.asarray(nlm0, dtype=np.complex128)[None, :]
    for j in range(Nt - 1):
        cur = step_arr(cur, (np.asarray(D[j]) + np.asarray(W[j]))[None], dt=dt, iota=iota, zeta=0.0, terms=("lrot",))
        out[j + 1] = cur[0]
    return out


# ------------------------------------------------------------------------------------------
# bounds and reduced form, reference: src/specfabpy.f90:764-771, 1039-1061
# ------------------------------------------------------------------------------------------

def apply_bounds_arr(nlm):
    n = _need_init()
    x = _farr(nlm, np.complex128, (n,))
    N = x.shape[0]
    out = np.empty((N, n), dtype=np.complex128, order="F")
    _lib.check(_lib.load().sfb_apply_bounds_arr(x.ctypes.data, out.ctypes.data, N, N))
    return out


def pin_array(a):
    """Page-lock an existing numpy array in place (cudaHostRegister) so that the host-array entry points (step_arr, ...)
    move it by DMA at PCIe speed; pageable arrays are ~4x slower.  Keep the array alive and call unpin_array(a) before it
    is freed.  Returns a."""
    if not isinstance(a, np.ndarray) or a.nbytes == 0:
        raise ValueError("pin_array needs a non-empty numpy array")
    if not (a.flags.c_contiguous or a.flags.f_contiguous):
        raise ValueError("pin_array needs a contiguous array")
    _lib.check(_lib.load().sfb_host_register(a.ctypes.data, a.nbytes))
    return a


def unpin_array(a):
    _lib.check(_lib.load().sfb_host_unregister(a.ctypes.data))


def apply_bounds_rnlm_arr_dev(rnlm, out=None):
    """apply_bounds on a resident reduced-form field: rnlm (rnlm_len, N) complex128 CUDA tensor, in place by default
    (what src/specfabpy/fenics/CPO.py:339-365 does per node; reference procedure src/dynamics.f90:530-557)."""
    import torch
    _need_init()
    if rnlm.dtype != torch.complex128 or not rnlm.is_cuda or not rnlm.is_contiguous() or rnlm.shape[0] != rnlm_len():
        raise ValueError("rnlm must be a contiguous CUDA complex128 tensor of shape (rnlm_len, N)")
    out = rnlm if out is None else out
    N = rnlm.shape[1]
    _lib.check(_lib.load().sfb_apply_bounds_rnlm_arr_dev(rnlm.data_ptr(), out.data_ptr(), N, N, N, _stream_ptr()))
    return out


def apply_bounds(nlm):
    """reference: src/specfabpy.f90:764-771"""
    return np.ascontiguousarray(apply_bounds_arr(np.asarray(nlm)[None, :])[0])


def rnlm_len():
    return _lib.load().sfb_rnlm_len()


def nlm_to_rnlm_arr(nlm):
    n = _need_init()
    x = _farr(nlm, np.complex128, (n,))
    N = x.shape[0]
    out = np.empty((N, rnlm_len()), dtype=np.complex128, order="F")
    _lib.check(_lib.load().sfb_nlm_to_rnlm_arr(x.ctypes.data, out.ctypes.data, N))
    return out


def rnlm_to_nlm_arr(rnlm):
    n = _need_init()
    x = _farr(rnlm, np.complex128, (rnlm_len(),))
    N = x.shape[0]
    out = np.empty((N, n), dtype=np.complex128, order="F")
    _lib.check(_lib.load().sfb_rnlm_to_nlm_arr(x.ctypes.data, out.ctypes.data, N))
    return out


def nlm_to_rnlm(nlm, rnlm_len_=None):
    """reference: src/specfabpy.f90:1046-1053"""
    return np.ascontiguousarray(nlm_to_rnlm_arr(np.asarray(nlm)[None, :])[0])


def rnlm_to_nlm(rnlm, nlm_len_=None):
    """reference: src/specfabpy.f90:1055-1061"""
    return np.ascontiguousarray(rnlm_to_nlm_arr(np.asarray(rnlm)[None, :])[0])


# ------------------------------------------------------------------------------------------
# structure tensors, eigenframe, enhancement factors (host arrays)
# ------------------------------------------------------------------------------------------

def _nlm15(nlm, k=15):
    """(N, >=k) complex -> Fortran-ordered (N, k) complex128 (only the first k coefficients are read: l<=4 for k=15,
    l<=6 for 28, l<=8 for 45)"""
    a = np.asarray(nlm, dtype=np.complex128)
    if a.ndim != 2 or a.shape[1] < k:
        raise ValueError("expected nlm of shape (N, nlm_len>=%d), got %s" % (k, a.shape))
    return np.asfortranarray(a[:, :k])


def a2_arr(nlm):
    """a2 of every node: (N,nlm_len) -> (N,3,3)          reference per node: src/specfabpy.f90:583-590"""
    _need_init()
    x = _nlm15(nlm)
    N = x.shape[0]
    out = np.empty((N, 3, 3), dtype=np.float64, order="F")
    _lib.check(_lib.load().sfb_a2_arr(x.ctypes.data, N, N, out.ctypes.data))
    return out


def a4_arr(nlm):
    """a4 of every node: (N,nlm_len) -> (N,3,3,3,3)      reference per node: src/specfabpy.f90:592-599"""
    _need_init()
    x = _nlm15(nlm)
    N = x.shape[0]
    out = np.empty((N, 3, 3, 3, 3), dtype=np.float64, order="F")
    _lib.check(_lib.load().sfb_a4_arr(x.ctypes.data, N, N, out.ctypes.data))
    return out


def a6_arr(nlm):
    """a6 of every node: (N,nlm_len) -> (N,3,3,3,3,3,3), needs L >= 6      reference per node: src/specfabpy.f90:601-608"""
    _need_init()
    x = _nlm15(nlm, 28)
    N = x.shape[0]
    out = np.empty((N,) + (3,) * 6, dtype=np.float64, order="F")
    _lib.check(_lib.load().sfb_a6_arr(x.ctypes.data, N, N, out.ctypes.data))
    return out


def E_CAFFE_arr(nlm, eps, Emin, Emax, n_grain):
    """E_CAFFE_arr(nlm (N,nlm_len), eps (N,3,3), Emin, Emax, n_grain) -> E (N,)     reference: src/specfabpy.f90:543-554"""
    _need_init()
    x = _nlm15(nlm, 45 if int(n_grain) == 3 else 15)
    N = x.shape[0]
    e = _farr(eps, np.float64, (3, 3))
    if e.shape[0] != N:
        raise ValueError("eps must have one 3x3 tensor per node")
    out = np.empty(N, dtype=np.float64)
    _lib.check(_lib.load().sfb_E_CAFFE_arr(x.ctypes.data, N, N, e.ctypes.data, float(Emin), float(Emax), int(n_grain), out.ctypes.data))
    return out


def pfJ_arr(nlm, Lmax=None):
    """pole-figure J index of every node, truncated at Lmax (default L) -> (N,)     reference per node: src/specfabpy.f90:729-736"""
    _need_init()
    Lmax = _state["L"] if Lmax is None else int(Lmax)
    x = _nlm15(nlm, (Lmax + 1) * (Lmax + 2) // 2)
    N = x.shape[0]
    out = np.empty(N, dtype=np.float64)
    _lib.check(_lib.load().sfb_pfJ_arr(x.ctypes.data, N, N, Lmax, out.ctypes.data))
    return out


def _ai_to_nlm_arr(rank, a):
    x = _farr(a, np.float64, (3,) * rank)
    N = x.shape[0]
    out = np.empty((N, (rank + 1) * (rank + 2) // 2), dtype=np.complex128, order="F")
    _lib.check(_lib.load().sfb_ai_to_nlm_arr(rank, x.ctypes.data, N, out.ctypes.data))
    return out


def a2_to_nlm_arr(a2_):
    """a2 (N,3,3) -> nlm (N,6)                 reference per node: src/specfabpy.f90:619-627"""
    return _ai_to_nlm_arr(2, a2_)


def a4_to_nlm_arr(a4_):
    """a4 (N,3,3,3,3) -> nlm (N,15)            reference per node: src/specfabpy.f90:629-637"""
    return _ai_to_nlm_arr(4, a4_)


def a6_to_nlm_arr(a6_):
    """a6 (N,3,3,3,3,3,3) -> nlm (N,28)        reference per node: src/specfabpy.f90:639-647"""
    return _ai_to_nlm_arr(6, a6_)


def a2_to_nlm(a2_):
    return np.ascontiguousarray(a2_to_nlm_arr(np.asarray(a2_)[None])[0])


def a4_to_nlm(a4_):
    return np.ascontiguousarray(a4_to_nlm_arr(np.asarray(a4_)[None])[0])


def a6_to_nlm(a6_):
    return np.ascontiguousarray(a6_to_nlm_arr(np.asarray(a6_)[None])[0])


def eig_arr(nlm):
    """a2 eigenframe of every node -> (ei (N,3,3) with ei[p,i,:] the i-th eigenvector, lami (N,3)),
    largest eigenvalue first            reference per node: src/specfabpy.f90:312-320"""
    _need_init()
    x = _nlm15(nlm)
    N = x.shape[0]
    ei = np.empty((N, 3, 3), dtype=np.float64, order="F")
    lami = np.empty((N, 3), dtype=np.float64, order="F")
    _lib.check(_lib.load().sfb_eig_arr(x.ctypes.data, N, N, ei.ctypes.data, lami.ctypes.data))
    return ei, lami


def eigframe_arr(M, plane="ij"):
    """eigframe_arr(M (N,3,3), plane) -> (ei, lami)        reference: src/specfabpy.f90:333-344"""
    m = _farr(M, np.float64, (3, 3))
    N = m.shape[0]
    ei = np.empty((N, 3, 3), dtype=np.float64, order="F")
    lami = np.empty((N, 3), dtype=np.float64, order="F")
    _lib.check(_lib.load().sfb_eigframe_arr(m.ctypes.data, N, str(plane).encode(), ei.ctypes.data, lami.ctypes.data))
    return ei, lami


def Eij_tranisotropic_arr(nlm, e1, e2, e3, Eij_grain, alpha, n_grain, return_status=False):
    """Eij_tranisotropic_arr(nlm (N,nlm_len), e1,e2,e3 (N,3), Eij_grain(2), alpha, n_grain) -> Eij (N,6)
    reference: src/specfabpy.f90:474-486.  return_status adds the per-node SFB_ST_* flags.
    n_grain: 1, 3 (needs L >= 8) or -3."""
    _need_init()
    x = _nlm15(nlm, 45 if int(n_grain) == 3 else 15)
    N = x.shape[0]
    es = [_farr(e, np.float64, (3,)) for e in (e1, e2, e3)]
    if any(e.shape[0] != N for e in es):
        raise ValueError("e1, e2, e3 must have one row per node of nlm (%d)" % N)
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    if g.shape != (2,):
        raise ValueError("Eij_grain must have 2 entries (Emm, Emt)")
    out = np.empty((N, 6), dtype=np.float64, order="F")
    st = np.zeros(N, dtype=np.int32)
    _lib.check(_lib.load().sfb_Eij_tranisotropic_arr(x.ctypes.data, N, N, es[0].ctypes.data, es[1].ctypes.data, es[2].ctypes.data,
                                                     g.ctypes.data, float(alpha), int(n_grain), out.ctypes.data, st.ctypes.data))
    return (out, st) if return_status else out


def Evw_tranisotropic_arr(nlm, v, w, tau, Eij_grain, alpha, n_grain, return_status=False):
    """Evw_tranisotropic batched over nodes: nlm (N,nlm_len), v, w (N,3), tau (N,3,3) -> Evw (N,)
    reference (per node): src/specfabpy.f90:379-388, src/enhancementfactors.f90:47-69.  n_grain: 1 or -3."""
    _need_init()
    x = _nlm15(nlm, 15)
    N = x.shape[0]
    vv, ww, tt = _farr(v, np.float64, (3,)), _farr(w, np.float64, (3,)), _farr(tau, np.float64, (3, 3))
    if vv.shape[0] != N or ww.shape[0] != N or tt.shape[0] != N:
        raise ValueError("v, w, tau must have one entry per node of nlm (%d)" % N)
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    if g.shape != (2,):
        raise ValueError("Eij_grain must have 2 entries (Emm, Emt)")
    out = np.empty(N, dtype=np.float64)
    st = np.zeros(N, dtype=np.int32)
    _lib.check(_lib.load().sfb_Evw_tranisotropic_arr(x.ctypes.data, N, N, vv.ctypes.data, ww.ctypes.data, tt.ctypes.data, g.ctypes.data,
                                                     float(alpha), int(n_grain), out.ctypes.data, st.ctypes.data))
    return (out, st) if return_status else out


def Evw_tranisotropic(nlm, v, w, tau, Eij_grain, alpha, n_grain):
    """scalar form with the reference's f2py signature (src/specfabpy.f90:379)"""
    return float(Evw_tranisotropic_arr(np.asarray(nlm)[None, :], np.asarray(v, float)[None, :], np.asarray(w, float)[None, :],
                                       np.asarray(tau, float)[None, :, :], Eij_grain, alpha, n_grain)[0])


def Eij_orthotropic_arr(nlm_1, nlm_2, nlm_3, e1, e2, e3, Eij_grain, alpha, n_grain):
    """Eij_orthotropic_arr(nlm_1, nlm_2, nlm_3 (N,nlm_len), e1,e2,e3 (N,3), Eij_grain(6), alpha, n_grain) -> Eij (N,6)
    reference: src/specfabpy.f90:488-500.  Pass nlm_3 = 0*nlm_1 (or None) to derive the third axis from the first two."""
    _need_init()
    x1, x2 = _nlm15(nlm_1), _nlm15(nlm_2)
    x3 = _nlm15(nlm_3) if nlm_3 is not None else None
    N = x1.shape[0]
    if x2.shape[0] != N or (x3 is not None and x3.shape[0] != N):
        raise ValueError("nlm_1, nlm_2, nlm_3 must have the same number of nodes")
    es = [_farr(e, np.float64, (3,)) for e in (e1, e2, e3)]
    if any(e.shape[0] != N for e in es):
        raise ValueError("e1, e2, e3 must have one row per node of nlm (%d)" % N)
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    if g.shape != (6,):
        raise ValueError("Eij_grain must have 6 entries (Ebb, Enn, Evv, Env, Ebv, Enb)")
    out = np.empty((N, 6), dtype=np.float64, order="F")
    _lib.check(_lib.load().sfb_Eij_orthotropic_arr(x1.ctypes.data, x2.ctypes.data, x3.ctypes.data if x3 is not None else None, N, N,
                                                   es[0].ctypes.data, es[1].ctypes.data, es[2].ctypes.data,
                                                   g.ctypes.data, float(alpha), int(n_grain), out.ctypes.data))
    return out


def Eij_eigenframe_arr(nlm, Eij_grain, alpha, n_grain, return_frame=False, return_status=False):
    """Fused a2 -> eigenframe -> Eij_tranisotropic (eigenenhancements) of every node -> Eij (N,6)
    [, ei (N,3,3), lami (N,3)] [, status].  Batches src/specfabpy/fenics/enhancementfactor.py:101-128."""
    _need_init()
    x = _nlm15(nlm, 45 if int(n_grain) == 3 else 15)
    N = x.shape[0]
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    out = np.empty((N, 6), dtype=np.float64, order="F")
    ei = np.empty((N, 3, 3), dtype=np.float64, order="F")
    lami = np.empty((N, 3), dtype=np.float64, order="F")
    st = np.zeros(N, dtype=np.int32)
    _lib.check(_lib.load().sfb_Eij_eigenframe_arr(x.ctypes.data, N, N, g.ctypes.data, float(alpha), int(n_grain), out.ctypes.data,
                                                  ei.ctypes.data, lami.ctypes.data, st.ctypes.data))
    res = (out,)
    if return_frame:
        res += (ei, lami)
    if return_status:
        res += (st,)
    return res if len(res) > 1 else out


# scalar (single-state) forms with the reference's exact signatures; they run the same kernels with N = 1
def a2(nlm):
    """reference: src/specfabpy.f90:583-590"""
    return np.ascontiguousarray(a2_arr(np.asarray(nlm)[None, :])[0])


def a4(nlm):
    """reference: src/specfabpy.f90:592-599"""
    return np.ascontiguousarray(a4_arr(np.asarray(nlm)[None, :])[0])


def a6(nlm):
    """reference: src/specfabpy.f90:601-608"""
    return np.ascontiguousarray(a6_arr(np.asarray(nlm)[None, :])[0])


def E_CAFFE(nlm, eps, Emin, Emax, n_grain):
    """reference: src/enhancementfactors.f90:301-331 (specfabpy E_CAFFE)"""
    return float(E_CAFFE_arr(np.asarray(nlm)[None, :], np.asarray(eps)[None, :, :], Emin, Emax, n_grain)[0])


def pfJ(nlm, Lmax=None):
    """reference: src/specfabpy.f90:729-736"""
    return float(pfJ_arr(np.asarray(nlm)[None, :], Lmax)[0])


def eig(nlm):
    """reference: src/specfabpy.f90:312-320 -> (ei[3,3], lami[3])"""
    ei, lami = eig_arr(np.asarray(nlm)[None, :])
    return np.ascontiguousarray(ei[0]), np.ascontiguousarray(lami[0])


def Eij_orthotropic(nlm_1, nlm_2, nlm_3, e1, e2, e3, Eij_grain, alpha, n_grain):
    """reference: src/specfabpy.f90:436-446 -> Eij[6]"""
    one = lambda v: None if v is None else np.asarray(v)[None, :]
    return np.ascontiguousarray(Eij_orthotropic_arr(one(nlm_1), one(nlm_2), one(nlm_3), one(e1), one(e2), one(e3), Eij_grain, alpha, n_grain)[0])


def Eij_tranisotropic(nlm, e1, e2, e3, Eij_grain, alpha, n_grain):
    """reference: src/specfabpy.f90:390-400 -> Eij[6]"""
    return np.ascontiguousarray(Eij_tranisotropic_arr(np.asarray(nlm)[None, :], np.asarray(e1)[None, :], np.asarray(e2)[None, :],
                                                      np.asarray(e3)[None, :], Eij_grain, alpha, n_grain)[0])


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
    return _step_dev(False, nlm, ugrad, tau, out, dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps)


def step_rnlm_arr_dev(rnlm, ugrad, tau=None, out=None, dt=0.0, iota=1.0, zeta=0.0, nu=1.0, Gamma0=0.0, Lambda=0.0,
                      terms=("lrot", "reg"), scheme="euler", nsteps=1):
    """step_arr_dev on reduced-form states: rnlm (rnlm_len, N) complex128 CUDA tensor (rows m >= 0, see step_rnlm_arr)."""
    return _step_dev(True, rnlm, ugrad, tau, out, dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps)


def _step_dev(reduced, nlm, ugrad, tau, out, dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps):
    import torch
    n = _need_init()
    lib = _lib.load()
    if reduced:
        n = rnlm_len()
    if nlm.dtype != torch.complex128 or not nlm.is_cuda or not nlm.is_contiguous() or nlm.shape[0] != n:
        raise ValueError("state must be a contiguous CUDA complex128 tensor of shape (%s, N)" % ("rnlm_len" if reduced else "nlm_len"))
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
    _lib.check((lib.sfb_step_rnlm_arr_dev if reduced else lib.sfb_step_arr_dev)(nlm.data_ptr(), out.data_ptr(), N, N, N, ugrad.data_ptr(), N,
                                    tau.data_ptr() if tau is not None else None, N, C.byref(o), _stream_ptr()))
    return out


def step_moments_Eij_rnlm_arr_dev(rnlm, ugrad, tau, Eij_grain, alpha, n_grain, out=None, dt=0.0, iota=1.0, zeta=0.0, nu=1.0,
                                  Gamma0=0.0, Lambda=0.0, terms=("lrot", "reg"), scheme="euler", nsteps=1, want_a2=False,
                                  want_frame=False, step=True):
    """step_moments_Eij_arr_dev for a resident field kept in REDUCED form: rnlm (rnlm_len, N) complex128 CUDA tensor (rows
    m >= 0, see step_rnlm_arr).  step=False: only a2 / eigenframe / Eij of the given state.  n_grain = 1 or -3.
    Returns a dict with 'rnlm', 'Eij' (6,N) and, when asked for, 'a2' (3,3,N), 'ei' (3,3,N), 'lami' (3,N)."""
    import torch
    _need_init()
    r = rnlm_len()
    if rnlm.dtype != torch.complex128 or not rnlm.is_cuda or not rnlm.is_contiguous() or rnlm.shape[0] != r:
        raise ValueError("rnlm must be a contiguous CUDA complex128 tensor of shape (rnlm_len, N)")
    N = rnlm.shape[1]
    out = rnlm if (out is None or not step) else out
    dev = rnlm.device
    res = {"rnlm": out, "Eij": torch.empty((6, N), dtype=torch.float64, device=dev)}
    if want_a2:
        res["a2"] = torch.empty((3, 3, N), dtype=torch.float64, device=dev)
    if want_frame:
        res["ei"] = torch.empty((3, 3, N), dtype=torch.float64, device=dev)
        res["lami"] = torch.empty((3, N), dtype=torch.float64, device=dev)
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    if g.shape != (2,):
        raise ValueError("Eij_grain must have 2 entries (Emm, Emt)")
    ptr = lambda k: res[k].data_ptr() if k in res else None
    lib = _lib.load()
    if step:
        o = _opts(dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps, None, None)
        _lib.check(lib.sfb_step_moments_Eij_rnlm_arr_dev(rnlm.data_ptr(), out.data_ptr(), N, N, N, ugrad.data_ptr(), N,
                                                         tau.data_ptr() if tau is not None else None, N, C.byref(o),
                                                         g.ctypes.data, float(alpha), int(n_grain), res["Eij"].data_ptr(),
                                                         ptr("a2"), ptr("ei"), ptr("lami"), None, _stream_ptr()))
    else:
        _lib.check(lib.sfb_Eij_eigenframe_rnlm_arr_dev(rnlm.data_ptr(), N, N, g.ctypes.data, float(alpha), int(n_grain),
                                                       res["Eij"].data_ptr(), ptr("a2"), ptr("ei"), ptr("lami"), None, _stream_ptr()))
    return res


def step_moments_Eij_arr_dev(nlm, ugrad, tau, Eij_grain, alpha, n_grain, out=None, dt=0.0, iota=1.0, zeta=0.0, nu=1.0,
                             Gamma0=0.0, Lambda=0.0, terms=("lrot", "reg"), scheme="euler", nsteps=1, want_a2=False, want_a4=False,
                             want_frame=False):
    """One FE time step on a resident field: fused step, then a2 / a4 / a2 eigenframe / eigenenhancements of the new
    state (SURVEY 8b step_moments_Eij_arr, BASELINE config 5).  Scalars Gamma0 / Lambda.  Returns a dict with 'nlm',
    'Eij' (6,N) and, when asked for, 'a2' (3,3,N), 'a4' (3,3,3,3,N), 'ei' (3,3,N), 'lami' (3,N)."""
    import torch
    n = _need_init()
    if nlm.dtype != torch.complex128 or not nlm.is_cuda or not nlm.is_contiguous() or nlm.shape[0] != n:
        raise ValueError("nlm must be a contiguous CUDA complex128 tensor of shape (nlm_len, N)")
    N = nlm.shape[1]
    out = nlm if out is None else out
    dev = nlm.device
    res = {"nlm": out, "Eij": torch.empty((6, N), dtype=torch.float64, device=dev)}
    if want_a2:
        res["a2"] = torch.empty((3, 3, N), dtype=torch.float64, device=dev)
    if want_a4:
        res["a4"] = torch.empty((3, 3, 3, 3, N), dtype=torch.float64, device=dev)
    if want_frame:
        res["ei"] = torch.empty((3, 3, N), dtype=torch.float64, device=dev)
        res["lami"] = torch.empty((3, N), dtype=torch.float64, device=dev)
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    if g.shape != (2,):
        raise ValueError("Eij_grain must have 2 entries (Emm, Emt)")
    o = _opts(dt, iota, zeta, nu, Gamma0, Lambda, terms, scheme, nsteps, None, None)
    ptr = lambda k: res[k].data_ptr() if k in res else None
    _lib.check(_lib.load().sfb_step_moments_Eij_arr_dev(nlm.data_ptr(), out.data_ptr(), N, N, N, ugrad.data_ptr(), N,
                                                        tau.data_ptr() if tau is not None else None, N, C.byref(o),
                                                        g.ctypes.data, float(alpha), int(n_grain), res["Eij"].data_ptr(),
                                                        ptr("a2"), ptr("a4"), ptr("ei"), ptr("lami"), None, _stream_ptr()))
    return res


def a2_arr_dev(nlm, out=None):
    """a2 of every node from a resident state: nlm (nlm_len,N) complex128 CUDA -> (3,3,N) float64 (Fortran (N,3,3))."""
    import torch
    _need_init()
    N = nlm.shape[1]
    if out is None:
        out = torch.empty((3, 3, N), dtype=torch.float64, device=nlm.device)
    _lib.check(_lib.load().sfb_a2_arr_dev(nlm.data_ptr(), N, N, out.data_ptr(), _stream_ptr()))
    return out


def Eij_tranisotropic_arr_dev(nlm, e1, e2, e3, Eij_grain, alpha, n_grain, out=None, status=None):
    """Eij in given frames: nlm (nlm_len,N), e1/e2/e3 (3,N) float64 CUDA (Fortran (N,3)) -> (6,N)."""
    import torch
    _need_init()
    N = nlm.shape[1]
    if out is None:
        out = torch.empty((6, N), dtype=torch.float64, device=nlm.device)
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    _lib.check(_lib.load().sfb_Eij_tranisotropic_arr_dev(nlm.data_ptr(), N, N, e1.data_ptr(), e2.data_ptr(), e3.data_ptr(), g.ctypes.data,
                                                         float(alpha), int(n_grain), out.data_ptr(),
                                                         status.data_ptr() if status is not None else None, _stream_ptr()))
    return out


def Eij_orthotropic_arr_dev(nlm_1, nlm_2, nlm_3, e1, e2, e3, Eij_grain, alpha, n_grain, out=None):
    """Eij of orthotropic grains on resident states: nlm_i (nlm_len,N) complex128 CUDA (nlm_3 may be None),
    e1/e2/e3 (3,N) float64 CUDA -> (6,N)."""
    import torch
    _need_init()
    N = nlm_1.shape[1]
    if out is None:
        out = torch.empty((6, N), dtype=torch.float64, device=nlm_1.device)
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    if g.shape != (6,):
        raise ValueError("Eij_grain must have 6 entries (Ebb, Enn, Evv, Env, Ebv, Enb)")
    _lib.check(_lib.load().sfb_Eij_orthotropic_arr_dev(nlm_1.data_ptr(), N, nlm_2.data_ptr(), N,
                                                       nlm_3.data_ptr() if nlm_3 is not None else None, N, N,
                                                       e1.data_ptr(), e2.data_ptr(), e3.data_ptr(), g.ctypes.data,
                                                       float(alpha), int(n_grain), out.data_ptr(), _stream_ptr()))
    return out


def Eij_eigenframe_arr_dev(nlm, Eij_grain, alpha, n_grain, out=None, ei=None, lami=None, status=None):
    """Fused a2 -> eigenframe -> Eij on a resident state -> (6,N) float64 CUDA tensor."""
    import torch
    _need_init()
    N = nlm.shape[1]
    if out is None:
        out = torch.empty((6, N), dtype=torch.float64, device=nlm.device)
    g = np.ascontiguousarray(Eij_grain, dtype=np.float64)
    _lib.check(_lib.load().sfb_Eij_eigenframe_arr_dev(nlm.data_ptr(), N, N, g.ctypes.data, float(alpha), int(n_grain), out.data_ptr(),
                                                      ei.data_ptr() if ei is not None else None,
                                                      lami.data_ptr() if lami is not None else None,
                                                      status.data_ptr() if status is not None else None, _stream_ptr()))
    return out
