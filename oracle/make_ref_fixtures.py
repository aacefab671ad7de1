"""Golden vectors from the COMPILED reference (TEST INFRASTRUCTURE ONLY).

    oracle/build_ref.sh                  # needs gfortran + LAPACK + numpy.f2py -> oracle/_ref/specfabpy
    python oracle/make_ref_fixtures.py   # -> tests/golden/ref_compiled.npz

drives the reference's own f2py module (specfabpy.specfab, interface src/specfabpy.f90:144-1239) over seeded inputs and stores
inputs and outputs of every procedure on the hot path of SURVEY.md section 8a.  tests/test_oracle_golden.py picks the file up
when it exists and holds oracle/specfab_oracle.py to it (1e-13 relative, eigenvectors through their projectors), which pins
the oracle -- and through it every GPU parity test -- to the reference running for real: complex `qt**(2.0)`, `matmul`
summation order, LAPACK `dsyev` / `dposv` incl. the failed-factorisation fallback ("parity unpinned").

This container and the GPU image have no Fortran compiler, so the committed state is the recipe, not the file.
`--standin` runs the same case list through a shim of the numpy oracle that mimics the specfabpy signatures: it proves the
generator and the consuming test work end to end (tests/test_oracle_golden.py::test_ref_fixture_pipeline) and must never be
written to tests/golden/ref_compiled.npz (refused below).
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "..", "tests", "golden", "ref_compiled.npz")
LS = (4, 8, 12, 20)
GRAIN, ALPHA = (1.0, 1e3), 0.0125


def load_reference():
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    from specfabpy import specfab as sf          # the f2py module built by build_ref.sh
    return sf


class OracleStandin:
    """specfabpy-shaped view of oracle/specfab_oracle.py (argument order of src/specfabpy.f90)."""

    def __init__(self):
        sys.path.insert(0, HERE)
        import specfab_oracle as orc
        self.o = orc

    def init(self, L):
        self.o.init(L)
        lm = np.array(self.o.lm_list(L)).T
        return lm, lm.shape[1]

    def M_LROT(self, nlm, eps, omg, iota, zeta): return self.o.M_LROT(eps, omg, iota, zeta)
    def M_DDRX(self, nlm, tau): return self.o.M_DDRX(nlm, tau)
    def M_DDRX_src(self, nlm, tau): return self.o.M_DDRX_src(tau)
    def M_CDRX(self, nlm): return self.o.M_CDRX()
    def M_REG(self, nlm, eps): return self.o.M_REG(eps)
    def a2(self, nlm): return self.o.a2(nlm)
    def a4(self, nlm): return self.o.a4(nlm)
    def eig(self, nlm): return self.o.eig(nlm)
    def apply_bounds(self, nlm): return self.o.apply_bounds(nlm)
    def Eij_tranisotropic(self, nlm, e1, e2, e3, Eij_grain, alpha, n_grain):
        return self.o.Eij_tranisotropic(nlm, e1, e2, e3, Eij_grain, alpha, n_grain)
    def Evw_tranisotropic(self, nlm, v, w, tau, Eij_grain, alpha, n_grain):
        return self.o.Evw_tranisotropic(v, w, tau, nlm, Eij_grain, alpha, n_grain)
    def nlm_LROT(self, nlm0, dt, Nt, D, W, iota):
        out = np.zeros((Nt, len(nlm0)), dtype=np.complex128)
        v = np.array(nlm0, dtype=np.complex128)
        for t in range(Nt):
            out[t] = v                            # src/dynamics.f90:99-110: row t is the state BEFORE step t
            v = v + dt * (self.o.M_LROT(D[t], W[t], iota, 0.0) @ v)
        return out


def cases(L, n):
    """Seeded inputs (same generators as tests/util.py so that failures can be replayed there)."""
    sys.path.insert(0, os.path.join(HERE, "..", "tests"))
    from util import random_states, random_ugrad, random_tau
    x = random_states(L, 4, 9000 + L, True, decay=0.5)
    x[3] = random_states(L, 1, 9100 + L, False)[0]              # one general complex vector
    ug = random_ugrad(4, 9200 + L)
    tau = random_tau(4, 9300 + L)
    return x, ug, tau


def generate(sf):
    out = {}
    for L in LS:
        lm, n = sf.init(L)
        out["L%d_lm" % L] = np.asarray(lm)
        x, ug, tau = cases(L, n)
        D = (ug + ug.transpose(0, 2, 1)) / 2
        W = (ug - ug.transpose(0, 2, 1)) / 2
        out["L%d_nlm" % L], out["L%d_ugrad" % L], out["L%d_tau" % L] = x, ug, tau
        out["L%d_M_LROT" % L] = np.array([sf.M_LROT(x[p], D[p], W[p], 1.0, 0.0) for p in range(4)])
        out["L%d_M_LROT_zeta" % L] = np.array([sf.M_LROT(x[p], D[p], W[p], 0.7, 0.3) for p in range(4)])
        out["L%d_M_DDRX_src" % L] = np.array([sf.M_DDRX_src(x[p], tau[p]) for p in range(4)])
        out["L%d_M_DDRX" % L] = np.array([sf.M_DDRX(x[p], tau[p]) for p in range(4)])
        out["L%d_M_CDRX" % L] = np.asarray(sf.M_CDRX(x[0]))
        out["L%d_M_REG" % L] = np.array([sf.M_REG(x[p], D[p]) for p in range(4)])
        out["L%d_a2" % L] = np.array([sf.a2(x[p]) for p in range(4)])
        out["L%d_a4" % L] = np.array([sf.a4(x[p]) for p in range(4)])
        out["L%d_apply_bounds" % L] = np.array([sf.apply_bounds(4 * x[p]) for p in range(3)])
        eigs = [sf.eig(x[p]) for p in range(3)]
        out["L%d_eig_ei" % L] = np.array([e[0] for e in eigs])
        out["L%d_eig_lami" % L] = np.array([e[1] for e in eigs])
        e = np.eye(3)
        out["L%d_Eij" % L] = np.array([sf.Eij_tranisotropic(x[p], e[0], e[1], e[2], GRAIN, ALPHA, 1) for p in range(3)])
        out["L%d_Eij_eigframe" % L] = np.array([sf.Eij_tranisotropic(x[p], eigs[p][0][0], eigs[p][0][1], eigs[p][0][2], GRAIN, ALPHA, 1)
                                               for p in range(3)])
        v, w = np.array([1.0, 2.0, -0.5]) / np.linalg.norm([1.0, 2.0, -0.5]), np.array([2.0, -1.0, 0.0]) / np.sqrt(5.0)
        tvw = np.outer(v, w) + np.outer(w, v)
        out["L%d_Evw" % L] = np.array([sf.Evw_tranisotropic(x[p], v, w, tvw, GRAIN, ALPHA, 1) for p in range(3)])
        # 20 Euler steps of lattice rotation (src/dynamics.f90:99-110 via specfabpy nlm_LROT)
        Nt = 20
        x0 = np.zeros(n, dtype=np.complex128)
        x0[0] = 1 / np.sqrt(4 * np.pi)
        out["L%d_nlm_LROT" % L] = np.asarray(sf.nlm_LROT(x0, 0.05, Nt, np.tile(D[0], (Nt, 1, 1)), np.tile(W[0], (Nt, 1, 1)), 1.0))
    # the failed-dposv fallback of the Taylor homogenisation (src/homogenizations.f90:174-185): unphysical L = 8 states
    sys.path.insert(0, os.path.join(HERE, "..", "tests"))
    from util import random_states
    sf.init(8)
    xs = random_states(8, 12, 51, True, decay=1.0) * 4
    xs[:, 0] = 1 / np.sqrt(4 * np.pi)
    e = np.eye(3)
    out["fallback_nlm"] = xs
    res = []
    for p in range(12):
        try:
            res.append(np.asarray(sf.Eij_tranisotropic(xs[p], e[0], e[1], e[2], GRAIN, ALPHA, 1), dtype=np.float64))
        except Exception:                      # the oracle raises where the reference would `stop`
            res.append(np.full(6, np.nan))
    out["fallback_Eij"] = np.array(res)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=GOLDEN)
    ap.add_argument("--standin", action="store_true", help="use the numpy oracle behind specfabpy's signatures (pipeline self-test)")
    a = ap.parse_args()
    if a.standin and os.path.abspath(a.out) == os.path.abspath(GOLDEN):
        sys.exit("refusing to write oracle output to the golden path: ref_compiled.npz must come from the compiled reference")
    sf = OracleStandin() if a.standin else load_reference()
    data = generate(sf)
    data["source"] = np.array("oracle stand-in (NOT reference output)" if a.standin else "compiled reference (oracle/build_ref.sh)")
    np.savez_compressed(a.out, **data)
    print("wrote", a.out, len(data), "arrays")
