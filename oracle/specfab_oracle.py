"""CPU oracle: a numpy restatement of specfab's fabric-evolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in specfab_b200/ (the product) imports this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

The reference (nicholasmr/specfab, Fortran 90) cannot be compiled in this image (no Fortran
compiler anywhere), so this file restates the reference arithmetic procedure by
procedure, each citing the reference file:line it follows (paths relative to /root/reference).
Parity status: the reference ships no golden vectors; this oracle is
pinned by (1) tests/golden/*.npz -- outputs of a mechanical evaluation of the reference's own
generated formula text with Fortran kind semantics (tools/f90eval.py, tools/make_golden.py),
(2) the analytic invariants listed in SURVEY.md section 8c, (3) an independent C restatement
(oracle/specfab_oracle.c).  It is NOT pinned against output of the compiled reference:
"parity unpinned" w.r.t. a gfortran build, LAPACK eigenvector signs and the failed-Cholesky
fallback state.

Fortran kind semantics reproduced here: un-suffixed real literals and
their intrinsics are real(4); they are rounded to float32 *before* promotion to double.  The
Gaunt tables hold float32 values (src/include/gaunt__body.f90).
"""
import math
import os
import numpy as np

try:  # LAPACK entry points the reference calls (src/frames.f90:75, src/homogenizations.f90:174)
    from scipy.linalg.lapack import dsyev as _dsyev, dposv as _dposv
except Exception:  # pragma: no cover
    _dsyev = _dposv = None

f32 = np.float32
Pi = 3.141592653589793  # src/header.f90:12


def _r4(x):
    """double value of a real(4) quantity"""
    return float(f32(x))


# real(4) constants, evaluated in float32 left to right
SQRT3_F = _r4(np.sqrt(f32(3.0)))                 # sqrt(3.)      src/dynamics.f90:79
SQRT6_F = _r4(np.sqrt(f32(6.0)))                 # sqrt(6.)      src/dynamics.f90:80-81
SQRT56_F = _r4(np.sqrt(f32(5.0) / f32(6.0)))     # sqrt(5./6)    src/dynamics.f90:85-86
SQRT23_F = _r4(np.sqrt(f32(2.0) / f32(3.0)))     # sqrt(2./3)
SQRT32_F = _r4(np.sqrt(f32(3.0) / f32(2.0)))     # sqrt(3./2)
TWOTHIRDS_F = _r4(f32(2.0) / f32(3.0))           # 2./3          src/dynamics.f90:576
SQRT2_F = _r4(np.sqrt(f32(2.0)))                 # sqrt(2.)      src/dynamics.f90:592

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "specfab_b200", "data", "gaunt_L20.npz")

# ---------------------------------------------------------------------------------------------
# header.f90 / gaunt.f90 : index conventions and tables
# ---------------------------------------------------------------------------------------------

_state = {"L": None, "n": None, "tables": None}


def nlm_len(L):
    """src/header.f90:23  nlm_lenvec(L) = (L+1)(L+2)/2"""
    return (L + 1) * (L + 2) // 2


def lm_list(L):
    """src/header.f90:33  (l,m) of every coefficient: even l, m=-l..l"""
    return [(l, m) for l in range(0, L + 1, 2) for m in range(-l, l + 1)]


def load_tables():
    """src/gaunt.f90:14-17 set_gaunts(): dense GC, GCm, GC_m1, GC_p1 (231,231,15) real(8)
    arrays holding float32 values (src/include/gaunt__head.f90:2, gaunt__body.f90)."""
    if _state["tables"] is None:
        d = np.load(_DATA)
        T = {}
        for nm in ("GC", "GCm", "GC_m1", "GC_p1"):
            a = np.zeros((231, 231, 15), dtype=np.float64)
            a[d[nm + "_i"], d[nm + "_j"], d[nm + "_k"]] = d[nm + "_v"].astype(np.float64)
            T[nm] = a
        _state["tables"] = T
    return _state["tables"]


def init(L):
    """src/specfabpy.f90:150-160 init(L) -> (lm[2,n], nlm_len); src/specfab.f90:37-53.
    L restricted to even 4..20 (tables are L=20; regcalib has even L only)."""
    if L % 2 or L < 4 or L > 20:
        raise ValueError("L must be even, 4 <= L <= 20")
    _state["L"] = L
    _state["n"] = nlm_len(L)
    load_tables()
    return np.array(lm_list(L), dtype=np.int32).T.copy(), _state["n"]


def _L():
    if _state["L"] is None:
        raise RuntimeError("oracle.init(L) not called")
    return _state["L"], _state["n"]


# ---------------------------------------------------------------------------------------------
# dynamics.f90 : quadrics
# ---------------------------------------------------------------------------------------------

def quad_rr(M):
    """src/dynamics.f90:563-579.  Index 0..4 <-> m=-2..2."""
    fsq = math.sqrt(2 * Pi / 15)
    xx, yy, zz = M[0][0], M[1][1], M[2][2]
    xy, xz, yz = M[0][1], M[0][2], M[1][2]
    q = [None] * 5
    q[0] = fsq * complex(xx - yy, 2 * xy)
    q[1] = (2 * fsq) * complex(xz, yz)
    c0 = TWOTHIRDS_F * math.sqrt(Pi / 5)            # -2./3*sqrt(Pi/5): real(4) 2./3 promoted
    q[2] = complex(-(c0 * (xx + yy - 2 * zz)), 0.0)
    q[3] = -(2 * fsq) * complex(xz, -yz)
    q[4] = fsq * complex(xx - yy, -2 * xy)
    return q


def quad_tp(M):
    """src/dynamics.f90:581-593.  Index 0..2 <-> m=-1..1."""
    fsq1 = math.sqrt(2 * Pi / 3)
    xy, xz, yz = M[0][1], M[0][2], M[1][2]
    return [fsq1 * complex(yz, -xz), fsq1 * complex(SQRT2_F * xy, 0.0), fsq1 * complex(-yz, -xz)]


# ---------------------------------------------------------------------------------------------
# dynamics.f90 : operators
# ---------------------------------------------------------------------------------------------

def lrot_weights(eps, omg, iota, zeta):
    """src/dynamics.f90:71-91: the four 6-vectors g0, gz, gn, gp."""
    eps = np.asarray(eps, dtype=np.float64)
    omg = np.asarray(omg, dtype=np.float64)
    epssq = eps @ eps
    with np.errstate(all="ignore"):
        zetanorm = np.float64(zeta) / np.sqrt(epssq[0, 0] + epssq[1, 1] + epssq[2, 2])  # 0/0 -> NaN, as the reference
    qe = quad_rr(iota * eps + zetanorm * epssq)
    qo = quad_tp(omg)
    i = 1j
    z = 0j
    c6 = 6.0 / SQRT6_F   # i*6/sqrt(6.): (i*6) is complex(8), divided by the promoted real(4)
    g0_rot = [z, z, z, z, z, z]
    gz_rot = [-((i * SQRT3_F) * qo[1]), z, z, z, z, z]
    gn_rot = [-(complex(0.0, c6) * qo[0]), z, z, z, z, z]
    gp_rot = [+(complex(0.0, c6) * qo[2]), z, z, z, z, z]
    g0_Tay = [3 * v for v in (z, qe[0], qe[1], qe[2], qe[3], qe[4])]
    gz_Tay = [z, -qe[0], z, z, z, qe[4]]
    gn_Tay = [SQRT56_F * qe[1], z, qe[0], SQRT23_F * qe[1], SQRT32_F * qe[2], 2 * qe[3]]
    gp_Tay = [SQRT56_F * qe[3], 2 * qe[1], SQRT32_F * qe[2], SQRT23_F * qe[3], qe[4], z]
    add = lambda a, b: np.array([x + y for x, y in zip(a, b)], dtype=np.complex128)
    return add(g0_rot, g0_Tay), add(gz_rot, gz_Tay), add(gn_rot, gn_Tay), add(gp_rot, gp_Tay)


def M_LROT(eps, omg, iota, zeta):
    """src/dynamics.f90:52-97 (lattice rotation operator, n x n complex)."""
    L, n = _L()
    T = load_tables()
    g0, gz, gn, gp = lrot_weights(eps, omg, iota, zeta)
    M = -1 * (T["GC"][:n, :n, :6] @ g0 + T["GCm"][:n, :n, :6] @ gz + T["GC_m1"][:n, :n, :6] @ gn + T["GC_p1"][:n, :n, :6] @ gp)
    return M


def doubleinner22(A, B):
    """src/tensorproducts.f90:153-159: A_ij B_ji"""
    A = np.asarray(A); B = np.asarray(B)
    return float(sum(sum(A[i, j] * B[j, i] for j in range(3)) for i in range(3)))


def ddrx_weights_raw(qt):
    """src/include/ddrx-coupling-weights.f90:1-16 -> (k, g[15]) before normalisation.
    qt**(2.0) (complex**real, lowered to cpow by gfortran) is evaluated as qt*qt."""
    qm2, qm1, q0, qp1, qp2 = qt
    s5 = _r4(np.sqrt(f32(5.0))); s15 = _r4(np.sqrt(f32(1.5))); s6 = _r4(np.sqrt(f32(6.0)))
    s14 = _r4(np.sqrt(f32(14.0))); s7 = _r4(np.sqrt(f32(7.0))); s2 = _r4(np.sqrt(f32(2.0))); s3 = _r4(np.sqrt(f32(3.0)))
    ms6 = _r4(f32(-1.0) * np.sqrt(f32(6.0)))          # (-1.0)*Sqrt((6.0))
    c2s14 = _r4(f32(2) * np.sqrt(f32(14.0)))          # 2*Sqrt((14.0)) (sign applied to the whole term)
    c4s7 = _r4(f32(4) * np.sqrt(f32(7.0)))            # 4*Sqrt((7.0))
    c3s5 = _r4(f32(3.0) * np.sqrt(f32(5.0)))          # 3.*Sqrt((5.0))
    k = (3 * math.sqrt(5 / Pi)) / 28.0
    g = [None] * 15
    g[0] = (7.0 * (q0 * q0 + (-2.0 * qm1) * qp1 + (2.0 * qm2) * qp2)) / s5
    g[1] = s15 * (qm1 * qm1) + (-2.0 * q0) * qm2
    g[2] = q0 * qm1 + (ms6 * qp1) * qm2
    g[3] = q0 * q0 + (-1.0 * qm1) * qp1 + (-2.0 * qm2) * qp2
    g[4] = q0 * qp1 + (ms6 * qm1) * qp2
    g[5] = s15 * (qp1 * qp1) + (-2.0 * q0) * qp2
    g[6] = (-(c2s14 * (qm2 * qm2))) / 3.0
    g[7] = (-((c4s7 * qm1) * qm2)) / 3.0
    g[8] = (-(4 * (s2 * (qm1 * qm1) + (s3 * q0) * qm2))) / 3.0
    g[9] = (-(4 * ((s6 * q0) * qm1 + qp1 * qm2))) / 3.0
    g[10] = (-(4 * (3.0 * (q0 * q0) + (4.0 * qm1) * qp1 + qm2 * qp2))) / c3s5
    g[11] = (-(4 * ((s6 * q0) * qp1 + qm1 * qp2))) / 3.0
    g[12] = (-(4 * (s2 * (qp1 * qp1) + (s3 * q0) * qp2))) / 3.0
    g[13] = (-((c4s7 * qp1) * qp2)) / 3.0
    g[14] = (-(c2s14 * (qp2 * qp2))) / 3.0
    return k, np.array(g, dtype=np.complex128)


def ddrx_weights(tau):
    """src/dynamics.f90:291-293: g = k*g * 5/doubleinner22(tau,tau) with qt = quad_rr(tau)."""
    tau = np.asarray(tau, dtype=np.float64)
    k, g = ddrx_weights_raw(quad_rr(tau))
    dd = doubleinner22(tau, tau)
    with np.errstate(all="ignore"):
        return np.array([((k * x) * 5) / dd for x in g], dtype=np.complex128)


def M_DDRX_src(tau):
    """src/dynamics.f90:277-298"""
    L, n = _L()
    return load_tables()["GC"][:n, :n, :15] @ ddrx_weights(tau)


def ev_D2(nlm, tau):
    """src/dynamics.f90:402-422: <D> = 5[(tau.tau):a2 - tau:a4:tau]/(tau:tau)"""
    tau = np.asarray(tau, dtype=np.float64)
    tauv = mat_to_vec(tau)
    tausq = tau @ tau
    norm = tausq[0, 0] + tausq[1, 1] + tausq[2, 2]
    a2v, a4v = f_ev_ck_Mandel(nlm)
    D = float(np.dot(mat_to_vec(tausq), a2v)) - float(np.dot(tauv, a4v @ tauv))
    with np.errstate(all="ignore"):
        return 5 * D / norm


def M_DDRX(nlm, tau):
    """src/dynamics.f90:251-275: M_DDRX_src - <D> I   (caller multiplies by Gamma0)"""
    M = M_DDRX_src(tau)
    Davg = ev_D2(nlm, tau)
    idx = np.arange(M.shape[0])
    M[idx, idx] = M[idx, idx] - Davg
    return M


def Ldiag(L):
    """src/dynamics.f90:22: -l(l+1) per coefficient"""
    return np.array([-(l * (l + 1)) for (l, m) in lm_list(L)], dtype=np.float64)


def M_CDRX():
    """src/dynamics.f90:474-492 (caller multiplies by Lambda)"""
    L, n = _L()
    return np.diag(Ldiag(L))


REGCALIB = {  # src/include/regcalib.f90:1-36  (expo, nu)
    4: (1.700, 1.9879322126397958), 6: (1.150, 3.0011508426238862), 8: (1.600, 5.7498069921352384),
    10: (2.000, 10.7048905312159288), 12: (2.000, 10.6068117205577668), 14: (2.000, 13.3591023418822363),
    16: (2.500, 15.3094482670021108), 18: (2.500, 16.4844589176829217), 20: (3.000, 19.9467342880730136),
}


def reg_diag(L):
    """src/dynamics.f90:510-514: diag of M0_REG = abs(Ldiag/(L(L+1)))**expo"""
    expo, nu = REGCALIB[L]
    return np.array([math.pow(abs(x / (L * (L + 1))), expo) for x in Ldiag(L)], dtype=np.float64)


def M_REG(D):
    """src/dynamics.f90:494-518: -nu*||D||_F * M0_REG"""
    L, n = _L()
    expo, nu = REGCALIB[L]
    ratemag = nu * float(np.sqrt(np.sum(np.asarray(D, dtype=np.float64) ** 2)))
    return np.diag(-ratemag * reg_diag(L))


def Sl(nlm, l):
    """src/idealstate.f90:80-93 power spectrum at degree l"""
    i0 = l * (l - 1) // 2
    blk = np.asarray(nlm)[i0:i0 + 2 * l + 1]
    return 1.0 / (2 * l + 1) * float(np.sum(np.abs(blk) ** 2))


# ---------------------------------------------------------------------------------------------
# mandel.f90
# ---------------------------------------------------------------------------------------------

_S2 = math.sqrt(2.0)


def mat_to_vec(M):
    """src/mandel.f90:15-24"""
    return np.array([M[0][0], M[1][1], M[2][2], _S2 * M[1][2], _S2 * M[0][2], _S2 * M[0][1]], dtype=np.float64)


def vec_to_mat(v):
    """src/mandel.f90:26-37"""
    s = _S2
    return np.array([[v[0], v[5] / s, v[4] / s], [v[5] / s, v[1], v[3] / s], [v[4] / s, v[3] / s, v[2]]], dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# moments.f90
# ---------------------------------------------------------------------------------------------

def cdiv(a, b):
    """complex(8) division as GCC expands it under -fcx-fortran-rules (Smith, true divisions)."""
    ar, ai, br, bi = a.real, a.imag, b.real, b.imag
    with np.errstate(all="ignore"):
        ar, ai, br, bi = np.float64(ar), np.float64(ai), np.float64(br), np.float64(bi)
        if abs(br) < abs(bi):
            ratio = br / bi
            div = (br * ratio) + bi
            tr = (ar * ratio) + ai
            ti = (ai * ratio) - ar
        else:
            ratio = bi / br
            div = (bi * ratio) + br
            tr = (ai * ratio) + ar
            ti = ai - (ar * ratio)
        return complex(tr / div, ti / div)


def decompose_nlm(nlm):
    """src/moments.f90:341-355 (0-based slices)"""
    nlm = np.asarray(nlm, dtype=np.complex128)
    n00 = complex(nlm[0])
    n2m = [complex(x) for x in nlm[1:6]]
    n4m = [complex(x) for x in nlm[6:15]] if nlm.size >= 15 else [0j] * 9
    return n00, n2m, n4m


def f_ev_c0(n00):
    """src/moments.f90:184-189"""
    return math.sqrt(4 * Pi) * n00.real


def f_ev_c2(n00, n2m):
    """src/moments.f90:191-200 + src/include/ev_c2__body.f90:1-17 (all d0 constants).
    n2m index 0..4 <-> m=-2..2; only m>=0 is used."""
    h0, h1, h2 = cdiv(n2m[2], n00), cdiv(n2m[3], n00), cdiv(n2m[4], n00)
    c = 0.5 * math.sqrt(2.0 / 3)
    ev = np.zeros((3, 3))
    ev[0, 0] = -(c * h0.real) + h2.real
    ev[1, 1] = -(c * h0.real) - h2.real
    ev[2, 2] = math.sqrt(2.0 / 3) * h0.real
    ev[0, 1] = ev[1, 0] = -h2.imag
    ev[0, 2] = ev[2, 0] = -h1.real
    ev[1, 2] = ev[2, 1] = +h1.imag
    return math.sqrt(2 / 15.0) * ev + np.eye(3) / 3.0


def f_ev_c4_Mandel(n00, n2m, n4m):
    """src/moments.f90:211-218 + src/include/ev_c4_Mandel__body.f90:1-37 (d0 constants; uses only
    the m>=0 coefficients).  n2m[2+m], n4m[4+m]."""
    s5, s6, s7, s10, s30, s70, s15, s3 = (math.sqrt(x) for x in (5.0, 6.0, 7.0, 10.0, 30.0, 70.0, 15.0, 3.0))
    r00 = n00.real
    r20, r21, r22 = n2m[2].real, n2m[3].real, n2m[4].real
    i21, i22 = n2m[3].imag, n2m[4].imag
    r40, r41, r42, r43, r44 = (n4m[4 + m].real for m in range(5))
    i41, i42, i43, i44 = (n4m[4 + m].imag for m in range(1, 5))
    k = (2 * math.sqrt(Pi)) / 105.0
    ev = np.zeros((6, 6))
    ev[0, 0] = 21.0 * r00 + (s5 * -6.0) * r20 + (s30 * 6.0) * r22 + 3.0 * r40 + (s10 * -2.0) * r42 + s70 * r44
    ev[0, 1] = 7.0 * r00 + (-2.0 * s5) * r20 + r40 + (-1.0 * s70) * r44
    ev[0, 2] = 7.0 * r00 + s5 * (r20 + s6 * r22) + -4.0 * r40 + (s10 * 2.0) * r42
    ev[0, 3] = s10 * (s6 * i21 + -1.0 * i41 + s7 * i43)
    ev[0, 4] = (s10 * -1.0) * ((3.0 * s6) * r21 + -3.0 * r41 + s7 * r43)
    ev[0, 5] = (-2.0 * s5) * (math.pow(3.0, 1.5) * i22 + -1.0 * i42 + s7 * i44)
    ev[1, 1] = 21.0 * r00 + (s5 * -6.0) * (r20 + s6 * r22) + 3.0 * r40 + s10 * (2.0 * r42 + s7 * r44)
    ev[1, 2] = 7.0 * r00 + s5 * (r20 + (-1.0 * s6) * r22) + -2.0 * (2.0 * r40 + s10 * r42)
    ev[1, 3] = s10 * ((3.0 * s6) * i21 + -3.0 * i41 + (-1.0 * s7) * i43)
    ev[1, 4] = (s15 * -2.0) * r21 + s10 * (r41 + s7 * r43)
    ev[1, 5] = (2.0 * s5) * ((-3.0 * s3) * i22 + i42 + s7 * i44)
    ev[2, 2] = 21.0 * r00 + (12.0 * s5) * r20 + 8.0 * r40
    ev[2, 3] = s10 * ((3.0 * s6) * i21 + 4.0 * i41)
    ev[2, 4] = (s10 * -1.0) * ((3.0 * s6) * r21 + 4.0 * r41)
    ev[2, 5] = (-2.0 * s5) * (s3 * i22 + 2.0 * i42)
    ev[3, 3] = 2.0 * (7.0 * r00 + s5 * (r20 + (-1.0 * s6) * r22) + -2.0 * (2.0 * r40 + s10 * r42))
    ev[3, 4] = (s10 * -2.0) * (s3 * i22 + 2.0 * i42)
    ev[3, 5] = s5 * ((-2.0 * s6) * r21 + 2.0 * (r41 + s7 * r43))
    ev[4, 4] = 2.0 * (7.0 * r00 + s5 * (r20 + s6 * r22) + -4.0 * r40 + (s10 * 2.0) * r42)
    ev[4, 5] = (2.0 * s5) * (s6 * i21 + -1.0 * i41 + s7 * i43)
    ev[5, 5] = 2.0 * (7.0 * r00 + (-2.0 * s5) * r20 + r40 + (-1.0 * s70) * r44)
    for a in range(6):
        for b in range(a):
            ev[a, b] = ev[b, a]
    return ev * k / f_ev_c0(n00)


def f_ev_ck_Mandel(nlm):
    """src/moments.f90:164-178 -> (a2v[6], a4v[6,6])"""
    n00, n2m, n4m = decompose_nlm(nlm)
    return mat_to_vec(f_ev_c2(n00, n2m)), f_ev_c4_Mandel(n00, n2m, n4m)


def a2(nlm):
    """src/moments.f90:37-44"""
    n00, n2m, n4m = decompose_nlm(nlm)
    return f_ev_c2(n00, n2m)


_A4_FILL = None


def f_ev_c4(n00, n2m, n4m):
    """src/moments.f90:202-209 + src/include/ev_c4__body.f90:1-94 (real(4) constants, both +-m
    coefficients, REAL() of a complex(8) sum).  Returns a4[3,3,3,3]."""
    S = lambda x: _r4(np.sqrt(f32(x)))
    s5, s30, s10, s70, s2, s3, s7, s6 = S(5.0), S(30.0), S(10.0), S(70.0), S(2.0), S(3.0), S(7.0), S(6.0)
    m12s5 = _r4(f32(-12.0) * np.sqrt(f32(5.0)))
    s30x6 = _r4(np.sqrt(f32(30.0)) * f32(6.0)); s30xm6 = _r4(np.sqrt(f32(30.0)) * f32(-6.0))
    s10x2 = _r4(np.sqrt(f32(10.0)) * f32(2.0)); s10xm2 = _r4(np.sqrt(f32(10.0)) * f32(-2.0))
    m3s3 = _r4(f32(-3.0) * np.sqrt(f32(3.0))); p3_15 = _r4(math.pow(3.0, 1.5))   # (3.0)**1.5: real(4) constant folded by gfortran (correctly rounded)
    ms7 = _r4(f32(-1.0) * np.sqrt(f32(7.0))); ms70 = _r4(f32(-1.0) * np.sqrt(f32(70.0)))
    p3s6 = _r4(f32(3.0) * np.sqrt(f32(6.0))); m3s6 = _r4(f32(-3.0) * np.sqrt(f32(6.0)))
    m4s5 = _r4(f32(-4.0) * np.sqrt(f32(5.0))); p2s5 = _r4(f32(2.0) * np.sqrt(f32(5.0)))
    ms6 = _r4(f32(-1.0) * np.sqrt(f32(6.0))); ms30 = _r4(f32(-1.0) * np.sqrt(f32(30.0)))
    ms3 = _r4(f32(-1.0) * np.sqrt(f32(3.0))); p12s5 = _r4(f32(12.0) * np.sqrt(f32(5.0)))
    i_s2 = complex(0.0, s2)      # (0,1)*Sqrt((2.0)) : complex(4) constant
    i_1 = complex(0.0, 1.0)
    n2 = lambda m: n2m[2 + m]
    n4 = lambda m: n4m[4 + m]
    k = math.sqrt(Pi / 5.0) / 21.0
    u = {}
    u[(1, 1, 1, 1)] = ((42.0 * n00 + m12s5 * n2(0) + s30x6 * n2(-2) + s30x6 * n2(+2) + 6.0 * n4(0) + s10xm2 * n4(-2) + s10xm2 * n4(+2) + s70 * n4(-4) + s70 * n4(+4)) / s5).real
    u[(1, 1, 1, 2)] = (i_s2 * (m3s3 * n2(-2) + p3_15 * n2(+2) + n4(-2) + -1.0 * n4(+2) + ms7 * n4(-4) + s7 * n4(+4))).real
    u[(1, 1, 1, 3)] = (p3s6 * n2(-1) + m3s6 * n2(+1) + -3.0 * n4(-1) + 3.0 * n4(+1) + s7 * n4(-3) + ms7 * n4(+3)).real
    u[(1, 1, 2, 2)] = ((14.0 * n00 + m4s5 * n2(0) + 2.0 * n4(0) + ms70 * n4(-4) + ms70 * n4(+4)) / s5).real
    u[(1, 1, 2, 3)] = (i_1 * (ms6 * n2(-1) + ms6 * n2(+1) + n4(-1) + n4(+1) + ms7 * n4(-3) + ms7 * n4(+3))).real
    u[(1, 1, 3, 3)] = ((14.0 * n00 + p2s5 * n2(0) + s30 * n2(-2) + s30 * n2(+2) + -8.0 * n4(0) + s10x2 * n4(-2) + s10x2 * n4(+2)) / s5).real
    u[(1, 2, 2, 2)] = (i_s2 * (m3s3 * n2(-2) + p3_15 * n2(+2) + n4(-2) + -1.0 * n4(+2) + s7 * n4(-4) + ms7 * n4(+4))).real
    u[(1, 2, 2, 3)] = (s6 * n2(-1) + ms6 * n2(+1) + -1.0 * n4(-1) + n4(+1) + ms7 * n4(-3) + s7 * n4(+3)).real
    u[(1, 2, 3, 3)] = (i_s2 * (ms3 * n2(-2) + s3 * n2(+2) + -2.0 * n4(-2) + 2.0 * n4(+2))).real
    u[(1, 3, 3, 3)] = (p3s6 * n2(-1) + m3s6 * n2(+1) + 4.0 * n4(-1) + -4.0 * n4(+1)).real
    u[(2, 2, 2, 2)] = ((42.0 * n00 + m12s5 * n2(0) + s30xm6 * n2(-2) + s30xm6 * n2(+2) + 6.0 * n4(0) + s10x2 * n4(-2) + s10x2 * n4(+2) + s70 * n4(-4) + s70 * n4(+4)) / s5).real
    u[(2, 2, 2, 3)] = (i_1 * (m3s6 * n2(-1) + m3s6 * n2(+1) + 3.0 * n4(-1) + 3.0 * n4(+1) + s7 * n4(-3) + s7 * n4(+3))).real
    u[(2, 2, 3, 3)] = ((14.0 * n00 + p2s5 * n2(0) + ms30 * n2(-2) + ms30 * n2(+2) + -8.0 * n4(0) + s10xm2 * n4(-2) + s10xm2 * n4(+2)) / s5).real
    u[(2, 3, 3, 3)] = (i_1 * (m3s6 * n2(-1) + m3s6 * n2(+1) + 4.0 * (-1.0 * n4(-1) + -1.0 * n4(+1)))).real
    u[(3, 3, 3, 3)] = ((2.0 * (21.0 * n00 + p12s5 * n2(0) + 8.0 * n4(0))) / s5).real
    ev = np.zeros((3, 3, 3, 3))
    for a in range(3):
        for b in range(3):
            for c in range(3):
                for d in range(3):
                    key = tuple(sorted((a + 1, b + 1, c + 1, d + 1)))
                    ev[a, b, c, d] = u[key]
    # Reference quirk reproduced on purpose: src/include/ev_c4__body.f90:78 reads
    # `ev(3,2,1,2)=ev(1,2,3,3)` (the symmetric alias would be ev(1,2,2,3)); a4() therefore
    # returns that single entry "wrong".  A drop-in must return what the reference returns.
    ev[2, 1, 0, 1] = u[(1, 2, 3, 3)]
    return ev * k / f_ev_c0(n00)


def a4(nlm):
    """src/moments.f90:46-55"""
    n00, n2m, n4m = decompose_nlm(nlm)
    return f_ev_c4(n00, n2m, n4m)


# ---------------------------------------------------------------------------------------------
# 6th / 8th order structure tensors (src/moments.f90:57-66,137-162,220-236)
# ---------------------------------------------------------------------------------------------
# ev_c6__body.f90 / ev_c8__body.f90 are 0.26 / 3.9 MB of generated linear forms with real(4) constants; like the
# orthotropic bodies they are not transcribed: tools/make_moment_tables.py interprets the reference text once and
# stores the coefficients of the canonical (sorted-index) entries -- the reference's 3^k separate assignments are
# exactly permutation symmetric (measured deviation 0, stored as *_permdev).  tests/golden pins them.

_HI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "specfab_b200", "data", "moments_l8.npz")
_hi_cache = {}


def _hi_tables():
    if not _hi_cache:
        import itertools
        d = np.load(_HI)
        for tag, rank in (("c6", 6), ("c8", 8)):
            uniq = {t: u for u, t in enumerate(itertools.combinations_with_replacement((0, 1, 2), rank))}
            full = np.zeros((3,) * rank, dtype=np.int64)
            for idx in itertools.product((0, 1, 2), repeat=rank):
                full[idx] = uniq[tuple(sorted(idx))]
            M = np.zeros((len(uniq), 45), dtype=np.complex128)
            M[d[tag + "_u"], d[tag + "_j"]] = d[tag + "_c"]
            _hi_cache[tag] = (M, float(d[tag + "_k"]), full)
    return _hi_cache


def _f_ev_hi(tag, nlm):
    M, k, full = _hi_tables()[tag]
    x = np.zeros(45, dtype=np.complex128)
    v = np.asarray(nlm, dtype=np.complex128)[:45]
    x[:v.size] = v
    u = (M @ x).real * k / f_ev_c0(complex(x[0]))          # ev = ev * k/f_ev_c0(n00)   src/moments.f90:226,235
    return u[full]


def a6(nlm):
    """src/moments.f90:57-66 + f_ev_c6 (:220-227)"""
    return _f_ev_hi("c6", nlm)


def a8(nlm):
    """f_ev_c8, src/moments.f90:229-236"""
    return _f_ev_hi("c8", nlm)


def f_ev_ck(nlm, opt="f"):
    """src/moments.f90:137-162 -> (ev_c2, ev_c4, ev_c6, ev_c8)"""
    n00, n2m, n4m = decompose_nlm(nlm)
    c2, c4 = f_ev_c2(n00, n2m), f_ev_c4(n00, n2m, n4m)
    if opt == "f":
        return c2, c4, a6(nlm), a8(nlm)
    return c2, c4, np.zeros((3,) * 6), np.zeros((3,) * 8)


def doubleinner42(A, B):
    """src/tensorproducts.f90:169-179: A_lkij B_ji"""
    return np.einsum("lkij,ji->lk", A, B)


def doubleinner44(A, B):
    """src/tensorproducts.f90:205-211: A_lkij B_jikl"""
    return float(np.einsum("lkij,jikl->", A, B))


def doubleinner62(A, B):
    """src/tensorproducts.f90:213-227: A_nmlkij B_ji"""
    return np.einsum("nmlkij,ji->nmlk", A, B)


def doubleinner82(A, B):
    """src/tensorproducts.f90:229-247: A_ponmlkij B_ji"""
    return np.einsum("ponmlkij,ji->ponmlk", A, B)


def doubleinner84(A, B):
    """src/tensorproducts.f90:249-263: A_ponmlkij B_jikl"""
    return np.einsum("ponmlkij,jikl->ponm", A, B)


def outerprodmat2(A, B):
    """src/tensorproducts.f90:129-135: A_ij B_kl"""
    return np.einsum("ij,kl->ijkl", A, B)


def ev_D4(nlm, tau):
    """src/dynamics.f90:424-448: average basal-plane RSS to the fourth power"""
    tau = np.asarray(tau, dtype=np.float64)
    tausq = tau @ tau
    tautau = outerprodmat2(tau, tau)
    tausqtau = outerprodmat2(tausq, tau)
    norm = tausq[0, 0] + tausq[1, 1] + tausq[2, 2]
    c2, c4, c6, c8 = f_ev_ck(nlm, "f")
    D = doubleinner22(tausq, doubleinner42(c4, tausq))
    D = D + doubleinner44(tautau, doubleinner84(c8, tautau))
    D = D - 2 * doubleinner44(tausqtau, doubleinner62(c6, tau))
    with np.errstate(all="ignore"):
        return float(35 / 2.0 * np.float64(D) / np.float64(norm ** 2))


def ev_D(nlm, tau, pw):
    """src/dynamics.f90:380-400"""
    return ev_D2(nlm, tau) if pw == 2 else (ev_D4(nlm, tau) if pw == 4 else 0.0)


def E_CAFFE(nlm, eps, Emin, Emax, n_grain):
    """src/enhancementfactors.f90:301-331 (Placidi et al. 2010)"""
    Dmax = [0.0, 5 / 2.0, 0.0, 35 / 8.0]
    n_RSS = 4 if n_grain == 3 else 2
    D = ev_D(nlm, eps, n_RSS)
    Dmaxpow = math.pow(Dmax[n_RSS - 1], 4.0 / n_RSS)
    if D < 1:
        gam = (4.0 / n_RSS) / Dmaxpow * (Emax - 1) / (1 - Emin)
        with np.errstate(all="ignore"):
            return float(Emin + (1 - Emin) * np.power(np.float64(D), gam))     # D < 0 -> NaN like Fortran's real**real
    return ((Emax - 1) * math.pow(D, 4.0 / n_RSS) + Dmaxpow - Emax) / (Dmaxpow - 1)


def pfJ(nlm, Lmax=None):
    """src/idealstate.f90:111-124: pole-figure J index truncated at Lmax"""
    L, _ = _L()
    Lmax = L if Lmax is None else Lmax
    return 4 * Pi * sum((2 * l + 1) * Sl(nlm, l) for l in range(0, Lmax + 1, 2))


# ---------------------------------------------------------------------------------------------
# state ingest: a2 / a4 / a6 -> nlm (src/moments.f90:68-92)
# ---------------------------------------------------------------------------------------------
# affine maps extracted from include/a{2,4,6}_to_nlm__body.f90 by tools/make_ingest_tables.py (the a4 map composed
# with a4_to_mat, src/mandel.f90:52-66); pinned by tests/golden (numeric interpretation of the same text).

_ING = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "specfab_b200", "data", "ingest_l6.npz")
_ing_cache = {}


def _ai_to_nlm(tag, A, rank, nrow):
    if not _ing_cache:
        _ing_cache.update(np.load(_ING))
    A = np.asarray(A, dtype=np.float64)
    if A.shape != (3,) * rank:
        raise ValueError("expected a tensor of shape %s" % ((3,) * rank,))
    a = A.reshape(-1, order="F")
    out = np.array(_ing_cache[tag + "_c0"], dtype=np.complex128)
    np.add.at(out, _ing_cache[tag + "_row"], _ing_cache[tag + "_c"] * a[_ing_cache[tag + "_flat"]])
    return out


def a2_to_nlm(a2_):
    """src/moments.f90:68-74 -> nlm(1:6)"""
    return _ai_to_nlm("a2", a2_, 2, 6)


def a4_to_nlm(a4_):
    """src/moments.f90:76-84 -> nlm(1:15)"""
    return _ai_to_nlm("a4", a4_, 4, 15)


def a6_to_nlm(a6_):
    """src/moments.f90:86-92 -> nlm(1:28)"""
    return _ai_to_nlm("a6", a6_, 6, 28)


# ---------------------------------------------------------------------------------------------
# frames.f90
# ---------------------------------------------------------------------------------------------

def eig3(M):
    """src/frames.f90:62-80: dsyev('V','U'), largest eigenvalue first; returns (e1,e2,e3,eigvals)."""
    w, v, info = _dsyev(np.array(M, dtype=np.float64, order="F"), compute_v=1, lower=0)
    return v[:, 2].copy(), v[:, 1].copy(), v[:, 0].copy(), np.array([w[2], w[1], w[0]])


def eigframe(M, plane="ij"):
    """src/frames.f90:24-60 -> (ei[3,3] rows = eigenvectors, lami[3])"""
    e1, e2, e3, lami = eig3(M)
    ei = np.array([e1, e2, e3])
    if plane == "ij":
        sort = [0, 1, 2]
    else:
        if plane == "xy":
            k = 2
        elif plane == "xz":
            k = 1
        else:
            raise ValueError('eigframe(M, plane): plane not any of ij,xy,xz')
        Imax = int(np.argmax(np.abs(ei[:, k])))
        sort = {0: [1, 2, 0], 1: [0, 2, 1], 2: [0, 1, 2]}[Imax]
    return ei[sort, :], lami[sort]


def eig(nlm):
    """src/frames.f90:14-22"""
    return eigframe(a2(nlm), "ij")


# ---------------------------------------------------------------------------------------------
# rheologies.f90 / homogenizations.f90 / enhancementfactors.f90  (n'=1 branch)
# ---------------------------------------------------------------------------------------------

def rheo_params_tranisotropic(Eij, d, n, ef):
    """src/rheologies.f90:123-135"""
    ne = ef * 2 / (n + 1)
    cI = (math.pow(Eij[0], ne) - 1) / (d - 1)
    cM = (d * (math.pow(Eij[0], ne) + 1) - 2) / (d - 1) - 2 * math.pow(Eij[1], ne)
    cL = math.pow(Eij[1], ne) - 1
    return cI, cM, cL


def _sachs_eps(tau, I2, cA, cB, cC, ev_etac0, ev_etac2, a4tau):
    """last line of src/homogenizations.f90:115 / :218"""
    return ev_etac0 * tau - cA * doubleinner22(ev_etac2, tau) * np.eye(3) + cB * a4tau + cC * (tau @ ev_etac2 + ev_etac2 @ tau)


def _sachs_n3_terms(tau, cB, cC, c2, c4, c6, c8):
    """<eta c^k> terms of the n'=3 branch, src/homogenizations.f90:97-102 (also :212-216 with the isotropic tensors)"""
    I2 = doubleinner22(tau, tau)
    tausq = tau @ tau
    ev_etac0 = I2 * 1.0 + cB * doubleinner22(doubleinner42(c4, tau), tau) + 2 * cC * doubleinner22(c2, tausq)
    ev_etac2 = I2 * c2 + cB * doubleinner42(doubleinner62(c6, tau), tau) + 2 * cC * doubleinner42(c4, tausq)
    ev_etac4 = I2 * c4 + cB * doubleinner62(doubleinner82(c8, tau), tau) + 2 * cC * doubleinner62(c6, tausq)
    return ev_etac0, ev_etac2, doubleinner42(ev_etac4, tau)


def rheo_fwd_tranisotropic_sachshomo(tau, nlm, Eij_grain, n_grain):
    """src/homogenizations.f90:69-116 (n_grain = 1, 3 and -3)"""
    tau = np.asarray(tau, dtype=np.float64)
    cA, cB, cC = rheo_params_tranisotropic(Eij_grain, 3, float(n_grain), 1)
    I2 = doubleinner22(tau, tau)
    if n_grain == 1:
        a2v, a4v = f_ev_ck_Mandel(nlm)
        ev_etac0 = 1.0
        ev_etac2 = vec_to_mat(a2v)
        a4tau = vec_to_mat(a4v @ mat_to_vec(tau))
    elif n_grain == 3:
        if np.asarray(nlm).size < 45:
            raise ValueError("specfab error: Sachs homogenization with n'=3 requires L >= 8.")
        ev_etac0, ev_etac2, a4tau = _sachs_n3_terms(tau, cB, cC, *f_ev_ck(nlm, "f"))
    elif n_grain == -3:
        a2v, a4v = f_ev_ck_Mandel(nlm)
        ev_etac0 = I2 * 1.0
        ev_etac2 = I2 * vec_to_mat(a2v)
        a4tau = I2 * vec_to_mat(a4v @ mat_to_vec(tau))
    else:
        raise ValueError("specfab error: unsupported n'")
    return _sachs_eps(tau, I2, cA, cB, cC, ev_etac0, ev_etac2, a4tau)


def anticommutator_Mandel(a):
    """src/homogenizations.f90:238-258"""
    s = _S2
    return np.array([
        [2 * a[0, 0], 0.0, 0.0, 0.0, s * a[0, 2], s * a[0, 1]],
        [0.0, 2 * a[1, 1], 0.0, s * a[1, 2], 0.0, s * a[0, 1]],
        [0.0, 0.0, 2 * a[2, 2], s * a[1, 2], s * a[0, 2], 0.0],
        [0.0, s * a[1, 2], s * a[1, 2], a[1, 1] + a[2, 2], a[0, 1], a[0, 2]],
        [s * a[0, 2], 0.0, s * a[0, 2], a[0, 1], a[0, 0] + a[2, 2], a[1, 2]],
        [s * a[0, 1], s * a[0, 1], 0.0, a[0, 2], a[1, 2], a[0, 0] + a[1, 1]]], dtype=np.float64)


TIKHONOV_F = _r4(1e-6)  # 1e-6 is a real(4) literal, src/homogenizations.f90:178


def taylor_P(nlm, Eij_grain, n_grain):
    """P of src/homogenizations.f90:165-170"""
    cA, cB, cC = rheo_params_tranisotropic(Eij_grain, 3, float(n_grain), -1)
    a2v, a4v = f_ev_ck_Mandel(nlm)
    a2mat = vec_to_mat(a2v)
    Lm = anticommutator_Mandel(a2mat)
    idv = np.array([1.0, 1, 1, 0, 0, 0])
    return np.eye(6) - cA * np.outer(idv, a2v) + cB * a4v + cC * Lm


def rheo_fwd_tranisotropic_taylorhomo(tau, nlm, Eij_grain, n_grain, return_status=False):
    """src/homogenizations.f90:145-189: dposv('L') on P (lower triangle only), Tikhonov fallback
    on failure.  status: 0 ok, 1 fallback used, 2 failed (reference: `stop`)."""
    P = taylor_P(nlm, Eij_grain, n_grain)
    b = mat_to_vec(np.asarray(tau, dtype=np.float64)).reshape(6, 1)
    c, x, info = _dposv(np.array(P, order="F"), np.array(b, order="F"), lower=1)
    status = 0
    if info != 0:
        # LAPACK leaves P partially factorised; scipy returns that state in c.
        Pf = np.array(c)
        P_reg = Pf.T @ Pf + TIKHONOV_F * np.eye(6)
        b_reg = Pf.T @ b
        c2, x, info2 = _dposv(np.array(P_reg, order="F"), np.array(b_reg, order="F"), lower=1)
        status = 1 if info2 == 0 else 2
    eps = vec_to_mat(x[:, 0])
    return (eps, status) if return_status else eps


_iso_ck = {}


def _ev_ck_iso():
    """src/homogenizations.f90:58-62: a^(k) of the isotropic state through the same f_ev_ck (real(4) round-off included)"""
    if not _iso_ck:
        x = np.zeros(45, dtype=np.complex128); x[0] = 1 / math.sqrt(4 * Pi)
        _iso_ck["t"] = f_ev_ck(x, "f")
    return _iso_ck["t"]


def rheo_fwd_tranisotropic_sachshomo__isotropic(tau, Eij_grain, n_grain):
    """src/homogenizations.f90:191-222"""
    tau = np.asarray(tau, dtype=np.float64)
    cA, cB, cC = rheo_params_tranisotropic(Eij_grain, 3, float(n_grain), 1)
    I2 = doubleinner22(tau, tau)
    if n_grain == 1:
        return (1 + 2.0 / 15 * cB + 2.0 / 3 * cC) * tau
    if n_grain == -3:
        return I2 * tau
    if n_grain == 3:
        ev_etac0, ev_etac2, a4tau = _sachs_n3_terms(tau, cB, cC, *_ev_ck_iso())
        return _sachs_eps(tau, I2, cA, cB, cC, ev_etac0, ev_etac2, a4tau)
    raise ValueError("specfab error: unsupported n'")


def rheo_fwd_tranisotropic_taylorhomo__isotropic(tau, Eij_grain, n_grain):
    """src/homogenizations.f90:224-236"""
    cA, cB, cC = rheo_params_tranisotropic(Eij_grain, 3, float(n_grain), -1)
    return np.asarray(tau, dtype=np.float64) / (1 + 2.0 / 15 * cB + 2.0 / 3 * cC)


def tau_vv(v):
    """src/enhancementfactors.f90:398-405"""
    v = np.asarray(v, dtype=np.float64)
    return np.eye(3) / 3.0 - np.outer(v, v)


def tau_vw(v, w):
    """src/enhancementfactors.f90:407-413"""
    v = np.asarray(v, dtype=np.float64); w = np.asarray(w, dtype=np.float64)
    return np.outer(v, w) + np.outer(w, v)


def Evw_tranisotropic(v, w, tau, nlm, Eij_grain, alpha, n_grain, return_status=False):
    """src/enhancementfactors.f90:47-69"""
    vw = np.outer(v, w)
    with np.errstate(all="ignore"):
        Es = doubleinner22(rheo_fwd_tranisotropic_sachshomo(tau, nlm, Eij_grain, n_grain), vw) / \
            doubleinner22(rheo_fwd_tranisotropic_sachshomo__isotropic(tau, Eij_grain, n_grain), vw)
        et, st = rheo_fwd_tranisotropic_taylorhomo(tau, nlm, Eij_grain, n_grain, return_status=True)
        Et = doubleinner22(et, vw) / doubleinner22(rheo_fwd_tranisotropic_taylorhomo__isotropic(tau, Eij_grain, n_grain), vw)
    E = (1 - alpha) * Es + alpha * Et
    return (E, st) if return_status else E


def Eij_tranisotropic(nlm, e1, e2, e3, Eij_grain, alpha, n_grain, return_status=False):
    """src/enhancementfactors.f90:23-45 -> (E11,E22,E33,E23,E13,E12)"""
    args = [(e1, e1, tau_vv(e1)), (e2, e2, tau_vv(e2)), (e3, e3, tau_vv(e3)),
            (e2, e3, tau_vw(e2, e3)), (e1, e3, tau_vw(e1, e3)), (e1, e2, tau_vw(e1, e2))]
    out, stat = [], 0
    for v, w, t in args:
        E, st = Evw_tranisotropic(v, w, t, nlm, Eij_grain, alpha, n_grain, return_status=True)
        out.append(E); stat = max(stat, st)
    out = np.array(out)
    return (out, stat) if return_status else out


# ---------------------------------------------------------------------------------------------
# discrete (grain-ensemble) lattice rotation, the reference's own cross-check of M_LROT (src/dynamics.f90:112-137)
# ---------------------------------------------------------------------------------------------

def dri_LROT(ri, D, W, iota):
    """src/dynamics.f90:112-123: d r_i/dt = (W + iota (r r^T D - D r r^T)) r_i for every grain axis r_i (rows of ri)"""
    ri = np.asarray(ri, dtype=np.float64)
    out = np.empty_like(ri)
    for j, r in enumerate(ri):
        mm = np.outer(r, r)
        out[j] = (W + iota * (mm @ D - D @ mm)) @ r
    return out


def ri_LROT(ri0, dt, Nt, D, W, iota):
    """src/dynamics.f90:125-137: Euler steps with renormalisation; D, W (Nt,3,3); returns (Nt, ngrains, 3)"""
    ri = np.empty((Nt,) + np.shape(ri0))
    ri[0] = ri0
    for i in range(Nt - 1):
        new = ri[i] + dt * dri_LROT(ri[i], D[i], W[i], iota)
        ri[i + 1] = new / np.linalg.norm(new, axis=1)[:, None]
    return ri


# ---------------------------------------------------------------------------------------------
# reduced form (src/reducedform.f90)
# ---------------------------------------------------------------------------------------------

def rlm_list():
    """(l, m >= 0) of the reduced state vector, src/reducedform.f90:48"""
    L, _ = _L()
    return [(l, m) for l in range(0, L + 1, 2) for m in range(0, l + 1)]


def I_all():
    """0-based positions of the m >= 0 coefficients in nlm, src/reducedform.f90:35"""
    return np.array([l * (l + 1) // 2 + m for l, m in rlm_list()])


def nlm_to_rnlm(nlm):
    """src/reducedform.f90:172-182"""
    return np.asarray(nlm, dtype=np.complex128)[I_all()]


def rnlm_to_nlm(rnlm):
    """src/reducedform.f90:160-170: n_l^-m = (-1)^m conj(n_l^m)"""
    L, n = _L()
    out = np.zeros(n, dtype=np.complex128)
    for (l, m), v in zip(rlm_list(), np.asarray(rnlm, dtype=np.complex128)):
        out[l * (l + 1) // 2 + m] = v
        if m:
            out[l * (l + 1) // 2 - m] = (-1) ** m * np.conj(v)
    return out


def reduce_M(M):
    """src/reducedform.f90:76-120 -> (Mrr, Mri, Mir, Mii) with
    d(rnlm)/dt = Mrr Re(rnlm) + Mri Im(rnlm) + i (Mir Re(rnlm) + Mii Im(rnlm))"""
    M = np.asarray(M, dtype=np.complex128)
    ia = I_all()
    r = len(ia)
    Mrr = np.zeros((r, r)); Mri = np.zeros((r, r)); Mir = np.zeros((r, r)); Mii = np.zeros((r, r))
    for jj, (l, m) in enumerate(rlm_list()):
        jp = ia[jj]
        jn = jp - 2 * m
        Vp, Qp = M[ia, jp].real, M[ia, jp].imag
        Vn, Qn = M[ia, jn].real, M[ia, jn].imag
        s = 1 if m % 2 == 0 else -1
        if jp == jn:
            s = 0
        Mrr[:, jj] += Vp + s * Vn
        Mri[:, jj] += -Qp + s * Qn
        Mir[:, jj] += Qp + s * Qn
        Mii[:, jj] += Vp - s * Vn
    return Mrr, Mri, Mir, Mii


# ---------------------------------------------------------------------------------------------
# time stepping (src/dynamics.f90:108; src/specfabpy/integrator.py:73-77)
# ---------------------------------------------------------------------------------------------

def operator(nlm, ugrad, tau, iota=1.0, zeta=0.0, Gamma0=0.0, Lambda=0.0, nu_mult=1.0,
             use_lrot=True, use_ddrx=False, use_cdrx=False, use_reg=True):
    """M = M_LROT + Gamma0*M_DDRX + Lambda*M_CDRX + nu_mult*M_REG, assembled as
    src/specfabpy/integrator.py:55-77 does (D, W = sym/skew parts of ugrad)."""
    L, n = _L()
    ugrad = np.asarray(ugrad, dtype=np.float64)
    D = (ugrad + ugrad.T) / 2
    W = (ugrad - ugrad.T) / 2
    M = np.zeros((n, n), dtype=np.complex128)
    if use_lrot:
        M = M + M_LROT(D, W, iota, zeta)
    if use_ddrx:
        M = M + Gamma0 * M_DDRX(nlm, tau)
    if use_cdrx:
        M = M + Lambda * M_CDRX()
    if use_reg:
        M = M + nu_mult * M_REG(D)
    return M


def rhs(nlm, ugrad, tau, **kw):
    return operator(nlm, ugrad, tau, **kw) @ np.asarray(nlm, dtype=np.complex128)


def step_euler(nlm, dt, ugrad, tau=None, **kw):
    """nlm + dt*matmul(M, nlm)   (src/dynamics.f90:108)"""
    nlm = np.asarray(nlm, dtype=np.complex128)
    return nlm + dt * rhs(nlm, ugrad, tau, **kw)


def step_rk4(nlm, dt, ugrad, tau=None, **kw):
    """Classical RK4 over the oracle RHS.  RK4 does not exist in the reference (SURVEY.md 8a/a9:
    docs use scipy RK45); BASELINE config 2 asks for it, forcing held constant over the step."""
    nlm = np.asarray(nlm, dtype=np.complex128)
    k1 = rhs(nlm, ugrad, tau, **kw)
    k2 = rhs(nlm + (dt / 2) * k1, ugrad, tau, **kw)
    k3 = rhs(nlm + (dt / 2) * k2, ugrad, tau, **kw)
    k4 = rhs(nlm + dt * k3, ugrad, tau, **kw)
    return nlm + (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


def apply_bounds(nlm):
    """src/dynamics.f90:530-557"""
    nlm = np.array(nlm, dtype=np.complex128)
    out = nlm.copy()
    S0 = nlm[0].real ** 2
    S2_rel = Sl(nlm, 2) / S0
    S4_rel = Sl(nlm, 4) / S0
    if S2_rel > 1.0:
        out[1:6] = nlm[1:6] / math.sqrt(S2_rel)
    if S4_rel > 1.0:
        out[6:15] = nlm[6:15] / math.sqrt(S4_rel)
    return out


def nlm_LROT(nlm0, dt, Nt, D, W, iota):
    """src/dynamics.f90:99-110: Euler integrator of lattice rotation only (zeta=0)."""
    L, n = _L()
    out = np.zeros((Nt, n), dtype=np.complex128)
    out[0] = nlm0
    for j in range(Nt - 1):
        out[j + 1] = out[j] + dt * (M_LROT(D[j], W[j], iota, 0.0) @ out[j])
    return out


# ---------------------------------------------------------------------------------------------
# Orthotropic grains: Eij_orthotropic  (src/enhancementfactors.f90:134-189)
# ---------------------------------------------------------------------------------------------
# The four moment bodies are 80-200 KB of generated bilinear forms with real(4) constants; they are NOT
# transcribed by hand: tools/make_orthotropic_tables.py interprets the reference text once (symbolically,
# Fortran kind semantics) and stores the coefficient tensors in specfab_b200/data/orthotropic_l4.npz.
# tests/golden/refbodies.npz (numeric interpretation of the same text) pins them.

_ORTH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "specfab_b200", "data", "orthotropic_l4.npz")
_orth_cache = {}


def _orth_tables():
    if not _orth_cache:
        d = np.load(_ORTH)
        for tag in ("v2", "v4", "c2b2", "c2v2"):
            _orth_cache[tag] = {k: d["%s_%s" % (tag, k)] for k in ("e", "p", "q", "c", "norm_p", "norm_q", "norm_c", "k", "rank")}
    return _orth_cache


def _bilinear(tag, blm, nlm):
    """ev = REAL(sum C b_p n_q) * k / norm   (src/moments.f90:242-311 with the include bodies)"""
    T = _orth_tables()[tag]
    b = np.zeros(15, dtype=np.complex128); n = np.zeros(15, dtype=np.complex128)
    bl = np.asarray(blm, dtype=np.complex128)[:15]; nl = np.asarray(nlm, dtype=np.complex128)[:15]
    b[:bl.size] = bl; n[:nl.size] = nl
    rank = int(T["rank"])
    ev = np.zeros(3 ** rank)
    np.add.at(ev, T["e"], (T["c"] * b[T["p"]] * n[T["q"]]).real)
    norm = float(np.sum((T["norm_c"] * b[T["norm_p"]] * n[T["norm_q"]]).real))
    return (ev * float(T["k"]) / norm).reshape((3,) * rank, order="F")


def a2_orth(blm, nlm):
    """src/moments.f90:242-258 + include/ev_v2__body.f90"""
    return _bilinear("v2", blm, nlm)


def a4_orth(blm, nlm):
    """src/moments.f90:260-276 + include/ev_v4__body.f90"""
    return _bilinear("v4", blm, nlm)


def a4_joint(blm, nlm):
    """src/moments.f90:278-293 + include/ev_c2b2__body.f90"""
    return _bilinear("c2b2", blm, nlm)


def a4_jointcross(blm, nlm):
    """src/moments.f90:295-311 + include/ev_c2v2__body.f90"""
    return _bilinear("c2v2", blm, nlm)


def ai_orthotropic(q1, q2, q3):
    """src/moments.f90:357-384 -> (a2_i[3], a4_ii[3], a4_jk[3])"""
    a2_i = [a2(q1), a2(q2), None]
    a4_ii = [a4(q1), a4(q2), None]
    a4_jk = [None, None, a4_joint(q1, q2)]
    if np.asarray(q3)[0].real > _r4(1e-8):
        a2_i[2] = a2(q3); a4_ii[2] = a4(q3)
        a4_jk[0] = a4_joint(q2, q3); a4_jk[1] = a4_joint(q1, q3)
    else:
        a2_i[2] = a2_orth(q1, q2); a4_ii[2] = a4_orth(q1, q2)
        a4_jk[0] = a4_jointcross(q2, q1); a4_jk[1] = a4_jointcross(q1, q2)
    return a2_i, a4_ii, a4_jk


def rheo_params_orthotropic(Eij, n):
    """src/rheologies.f90:186-206 -> (lami[6], gam)"""
    B = [math.pow(x, 2 / (n + 1)) for x in Eij]
    lami = [-B[0] + B[1] + B[2], +B[0] - B[1] + B[2], +B[0] + B[1] - B[2], B[3], B[4], B[5]]
    gam = 2 * B[0] * B[1] + 2 * B[0] * B[2] + 2 * B[1] * B[2] - B[0] ** 2 - B[1] ** 2 - B[2] ** 2
    return lami, gam


def a4_sym2(a):
    """src/tensorproducts.f90:53-64: X(j,k,:,:) = (a(j,k,:,:) + a(:,:,j,k))/2"""
    return (a + a.transpose(2, 3, 0, 1)) / 2


def a4_sym4(a):
    """src/tensorproducts.f90:66-82: X(i,j,k,l) = (a(i,k,j,l) + a(k,j,i,l) + a(i,l,k,j) + a(l,j,k,i))/4"""
    return (a.transpose(0, 2, 1, 3) + a.transpose(2, 1, 0, 3) + a.transpose(0, 3, 2, 1) + a.transpose(3, 1, 2, 0)) / 4


def _check_sym4():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 3, 3, 3))
    X = a4_sym4(a)
    i, j, k, l = 0, 1, 2, 1
    assert abs(X[i, j, k, l] - (a[i, k, j, l] + a[k, j, i, l] + a[i, l, k, j] + a[l, j, k, i]) / 4) < 1e-15


def rheo_fwd_orthotropic_sachshomo(tau, a2_i, a4_ii, a4_jk, Eij_grain, n_grain):
    """src/homogenizations.f90:288-319"""
    lami, gam = rheo_params_orthotropic(Eij_grain, float(n_grain))
    ci = [4.0 / 3 * x for x in lami[:3]] + [2 * x for x in lami[3:]]
    ji, ki = [1, 2, 0], [2, 0, 1]
    M2 = [(a4_ii[ji[i]] + a4_ii[ki[i]] - 2 * a4_sym2(a4_jk[i])) / 4 for i in range(3)]
    H2 = [a4_sym4(a4_jk[i]) for i in range(3)]
    if n_grain != 1:
        return np.zeros((3, 3))           # "n_grain not supported, silently return 0"
    Q = ci[0] * M2[0] + ci[1] * M2[1] + ci[2] * M2[2] + ci[3] * H2[0] + ci[4] * H2[1] + ci[5] * H2[2]
    tau = np.asarray(tau, dtype=np.float64)
    # doubleinner42: eps(l,k) = sum_ij Q(l,k,i,j) tau(j,i)     src/tensorproducts.f90:171-181
    return np.einsum("lkij,ji->lk", Q, tau)


def Evw_orthotropic(v, w, tau, q1, q2, q3, Eij_grain, alpha, n_grain):
    """src/enhancementfactors.f90:158-189 (Sachs only)"""
    vw = np.outer(v, w)
    q1 = np.asarray(q1, dtype=np.complex128)
    qiso = np.zeros_like(q1); qiso[0] = q1[0]
    A = ai_orthotropic(q1, q2, q3)
    Aiso = ai_orthotropic(qiso, qiso, qiso)
    with np.errstate(all="ignore"):      # IEEE division like the compiled reference (0/0 -> NaN for n_grain /= 1)
        return float(np.float64(doubleinner22(rheo_fwd_orthotropic_sachshomo(tau, *A, Eij_grain, n_grain), vw)) /
                     np.float64(doubleinner22(rheo_fwd_orthotropic_sachshomo(tau, *Aiso, Eij_grain, n_grain), vw)))


def Eij_orthotropic(q1, q2, q3, e1, e2, e3, Eij_grain, alpha, n_grain):
    """src/enhancementfactors.f90:134-156 -> (E11,E22,E33,E23,E13,E12)"""
    args = [(e1, e1, tau_vv(e1)), (e2, e2, tau_vv(e2)), (e3, e3, tau_vv(e3)),
            (e2, e3, tau_vw(e2, e3)), (e1, e3, tau_vw(e1, e3)), (e1, e2, tau_vw(e1, e2))]
    return np.array([Evw_orthotropic(v, w, t, q1, q2, q3, Eij_grain, alpha, n_grain) for v, w, t in args])
