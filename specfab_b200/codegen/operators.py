"""Build-time only: turn the Gaunt table data into the fixed real operator coefficients the
CUDA kernels are generated from.

Reference algebra (paths relative to /root/reference):
  M_LROT  = -(GC.g0 + GCm.gz + GC_m1.gn + GC_p1.gp)                 src/dynamics.f90:78-96
  M_DDRX_src = sum_k GC(:,:,k) g_k,  k=1..15                         src/dynamics.f90:295-297
The weight vectors g0..gp are linear in the 8 quadric coefficients qe(-2:2), qo(-1:1)
(src/dynamics.f90:78-91), so  M_LROT(i,j) = qe[D]*A(i,j) + (i*qo[D])*B(i,j)  with D = m_i - m_j
and A, B real, fixed.  A and B are formed here in double from the float32-valued table entries
and the reference's real(4) constants; every product of two float32-valued doubles is exact in
double, so A/B differ from the reference's on-the-fly sums only by final roundings (<= 1 ulp).

Mirror symmetry used by the kernels (two lanes per node share one instruction stream):
  A(l_i,-m_i; l_j,-m_j) =  A(l_i,m_i; l_j,m_j),   B(mirror) = -B,   GC(mirror, k(lk,-mk)) = GC.
These identities hold EXACTLY for the reference's 6-digit tables; they are asserted below.
"""
import os
import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "gaunt_L20.npz")
f32 = np.float32


def _r4(x):
    return float(f32(x))


# real(4) constants of src/dynamics.f90:79-86 (same evaluation as SURVEY.md A.1)
SQRT3_F = _r4(np.sqrt(f32(3.0)))
C6 = 6.0 / _r4(np.sqrt(f32(6.0)))
S56 = _r4(np.sqrt(f32(5.0) / f32(6.0)))
S23 = _r4(np.sqrt(f32(2.0) / f32(3.0)))
S32 = _r4(np.sqrt(f32(3.0) / f32(2.0)))

CAT = [(l, m) for l in (0, 2, 4) for m in range(-l, l + 1)]        # catalyst (lk,mk), k = 0..14
CIDX = {c: k for k, c in enumerate(CAT)}

_cache = {}


def dense_tables():
    if "T" not in _cache:
        d = np.load(_DATA)
        T = {}
        for nm in ("GC", "GCm", "GC_m1", "GC_p1"):
            a = np.zeros((231, 231, 15), dtype=np.float64)
            a[d[nm + "_i"], d[nm + "_j"], d[nm + "_k"]] = d[nm + "_v"].astype(np.float64)
            T[nm] = a
        _cache["T"] = T
    return _cache["T"]


def lm_list(L):
    return [(l, m) for l in range(0, L + 1, 2) for m in range(-l, l + 1)]


def idx(l, m):
    return l * (l + 1) // 2 + m


class Operators:
    """Dense (n x n) real coefficient matrices for truncation L.
       A[D] (D=-2..2): coefficient of qe[D];  B[D] (D=-1..1): coefficient of i*qo[D];
       G[(lk,mk)]: GC(:,:,k) slices (DDRX)."""

    def __init__(self, L):
        T = dense_tables()
        n = (L + 1) * (L + 2) // 2
        self.L, self.n = L, n
        GC, GCm, Gm1, Gp1 = (T[k][:n, :n, :] for k in ("GC", "GCm", "GC_m1", "GC_p1"))
        A = {}
        A[-2] = -(3 * GC[:, :, 1] - GCm[:, :, 1] + Gm1[:, :, 2])
        A[-1] = -(3 * GC[:, :, 2] + S56 * Gm1[:, :, 0] + S23 * Gm1[:, :, 3] + 2 * Gp1[:, :, 1])
        A[0] = -(3 * GC[:, :, 3] + S32 * Gm1[:, :, 4] + S32 * Gp1[:, :, 2])
        A[1] = -(3 * GC[:, :, 4] + 2 * Gm1[:, :, 5] + S56 * Gp1[:, :, 0] + S23 * Gp1[:, :, 3])
        A[2] = -(3 * GC[:, :, 5] + GCm[:, :, 5] + Gp1[:, :, 4])
        B = {0: SQRT3_F * GCm[:, :, 0], -1: C6 * Gm1[:, :, 0], 1: -C6 * Gp1[:, :, 0]}
        G = {c: GC[:, :, k].copy() for k, c in enumerate(CAT)}
        self.A, self.B, self.G = A, B, G
        self.lm = lm_list(L)
        self._check()

    def _check(self):
        L, n, lm = self.L, self.n, self.lm
        mi = np.array([m for (l, m) in lm])
        dm = mi[:, None] - mi[None, :]
        mir = np.array([idx(l, -m) for (l, m) in lm])
        for D, a in self.A.items():
            assert np.all((a == 0) | (dm == D)), "A selection rule"
            assert np.array_equal(a[mir][:, mir], self.A[-D]), "A mirror symmetry"
        for D, b in self.B.items():
            assert np.all((b == 0) | (dm == D)), "B selection rule"
            assert np.array_equal(b[mir][:, mir], -self.B[-D]), "B mirror antisymmetry"
            li = np.array([l for (l, m) in lm])
            assert np.all((b == 0) | (li[:, None] == li[None, :])), "B couples equal l only"
        for (lk, mk), g in self.G.items():
            assert np.all((g == 0) | (dm == mk)), "GC selection rule"
            assert np.array_equal(g[mir][:, mir], self.G[(lk, -mk)]), "GC mirror symmetry"

    def dense_lrot(self, qe, qo):
        """Reassemble M_LROT from A/B (used by the CPU tests to validate the decomposition)."""
        M = np.zeros((self.n, self.n), dtype=np.complex128)
        for D in range(-2, 3):
            M += qe[D + 2] * self.A[D]
        for D in range(-1, 2):
            M += (1j * qo[D + 1]) * self.B[D]
        return M
