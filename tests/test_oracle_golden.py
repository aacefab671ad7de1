"""CPU: pins of the oracle.  The golden file holds outputs of the REFERENCE'S OWN
formula text interpreted with Fortran kind semantics (tools/make_golden.py); the oracle's hand
restatement must reproduce them to rounding."""
import os

import numpy as np
import pytest

import specfab_oracle as o

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "refbodies.npz"))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max()


def test_generated_bodies_match_reference_text():
    for c in range(G["nlm"].shape[0]):
        nlm = G["nlm"][c]
        n00, n2m, n4m = o.decompose_nlm(nlm)
        c0 = o.f_ev_c0(n00)
        assert rel(o.f_ev_c4(n00, n2m, n4m), G["a4_ev"][c] * G["a4_k"][c] / c0) < 1e-15       # ev_c4__body.f90 (real(4))
        assert rel(o.f_ev_c4_Mandel(n00, n2m, n4m), G["a4M_ev"][c] * G["a4M_k"][c] / c0) < 1e-15  # ev_c4_Mandel__body.f90
        assert rel(o.f_ev_c2(n00, n2m), np.sqrt(2 / 15.0) * G["c2_ev"][c] + np.eye(3) / 3) < 1e-15  # ev_c2__body.f90
        assert rel(np.array(o.quad_rr(G["sym"][c])), G["quad_rr"][c]) < 1e-15                 # dynamics.f90:571-577
        assert rel(np.array(o.quad_tp(G["skew"][c])), G["quad_tp"][c]) < 1e-15                # dynamics.f90:592
        assert rel(np.array(o.lrot_weights(G["sym"][c], G["skew"][c], 1.0, 0.0)), G["lrot_g"][c]) < 1e-15  # dynamics.f90:78-91
        k, g = o.ddrx_weights_raw(list(G["quad_rr"][c]))
        assert k == G["ddrx_k"][c] and rel(g, G["ddrx_g"][c]) < 5e-16                          # ddrx-coupling-weights.f90


def test_a4_reference_alias_quirk_is_in_the_golden_data():
    """src/include/ev_c4__body.f90:78 assigns ev(3,2,1,2)=ev(1,2,3,3): the reference's a4 is not fully symmetric."""
    ev = G["a4_ev"][0]
    assert ev[2, 1, 0, 1] == ev[0, 1, 2, 2] and ev[2, 1, 0, 1] != ev[0, 1, 1, 2]


def test_real4_constants():
    """SURVEY.md A.1 table"""
    assert o.SQRT3_F == 1.7320507764816284 and o.SQRT56_F == 0.9128709435462952
    assert o.SQRT23_F == 0.8164966106414795 and o.SQRT32_F == 1.2247449159622192
    assert o.TWOTHIRDS_F == 0.6666666865348816 and o.SQRT2_F == 1.4142135381698608
    assert 6.0 / o.SQRT6_F == 2.449489653641921 and o.TIKHONOV_F == 9.999999974752427e-07
    assert o.load_tables()["GC"][0, 0, 0] == 0.2820949852466583


def test_isotropic_and_trace_pins():
    L = 8
    lm, n = o.init(L)
    assert n == 45 and tuple(lm[:, 6]) == (4, -4)
    iso = np.zeros(n, complex); iso[0] = 1 / np.sqrt(4 * np.pi)
    assert np.array_equal(o.a2(iso), np.eye(3) / 3)
    a4iso = np.zeros((3, 3, 3, 3))
    I = np.eye(3)
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    a4iso[i, j, k, l] = (I[i, j] * I[k, l] + I[i, k] * I[j, l] + I[i, l] * I[j, k]) / 15
    assert 1e-10 < np.abs(o.a4(iso) - a4iso).max() < 1e-7      # only ~1e-8: real(4) constants
    e = np.eye(3)
    assert np.abs(o.Eij_tranisotropic(iso, e[0], e[1], e[2], (1, 1e3), 0.0125, 1) - 1).max() < 1e-12
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert abs(np.trace(o.a2(x)) - 1) < 1e-14


def test_single_maximum_enhancement_pin():
    """delta function, ice 'linear' -> E_mt = 9.97 ('=10'), E_mm = 0.00997"""
    o.init(8)
    d = np.zeros(45, complex)
    d[0] = 1 / np.sqrt(4 * np.pi); d[3] = np.sqrt(5 / (4 * np.pi)); d[10] = 3 / np.sqrt(4 * np.pi)
    e = np.eye(3)
    E = o.Eij_tranisotropic(d, e[0], e[1], e[2], (1, 1e3), 0.0125, 1)
    assert np.allclose(E, [9.97005242e-3] * 3 + [9.97005242, 9.97005242, 9.97005242e-3], rtol=1e-8)


def test_fabdyn_lrot_scenario_and_row0_residual():
    """docs/snippets/fabdyn-LROT.py scenario + the ~1e-6 first-row residual of M_LROT"""
    o.init(8)
    ug = np.diag([.5, .5, -1.])
    M = o.M_LROT(ug, np.zeros((3, 3)), 1.0, 0.0)
    assert abs(np.abs(M[0]).max() - 1.4436e-6) < 1e-9 and np.count_nonzero(M) == 100 and np.abs(M.imag).max() == 0
    x = np.zeros(45, complex); x[0] = 1 / np.sqrt(4 * np.pi)
    for _ in range(25):
        x = o.step_euler(x, 0.05, ug, use_reg=False)
    assert np.allclose(np.diag(o.a2(x)), [0.09693296, 0.09693296, 0.80613408], atol=1e-8)
    assert abs(x[3].real / x[0].real - 1.5858219082588) < 1e-12 and abs(x[10].real / x[0].real - 1.5872145434530) < 1e-12


def test_fabdyn_ddrx_scenario():
    o.init(8)
    S = np.diag([.5, .5, -1.])
    x = np.zeros(45, complex); x[0] = 1 / np.sqrt(4 * np.pi)
    assert o.ev_D2(x, S) == 0.9999999999999997
    n00 = x[0].real
    for _ in range(24):
        x = x + 0.05 * ((10 * o.M_DDRX(x, S)) @ x)
    assert np.allclose(np.diag(o.a2(x)), [0.26002527, 0.26002527, 0.47994946], atol=1e-8)
    assert abs(o.ev_D2(x, S) - 1.75353373) < 1e-8 and abs(x[0].real / n00 - 1 - 1.435e-5) < 1e-8


def test_lrot_preserves_reality_symmetry():
    """tests/reduced-form/reduced-form.f90 -- n_l^-m = (-1)^m conj(n_l^m) is preserved"""
    L = 8
    lm, n = o.init(L)
    rng = np.random.default_rng(0)
    ug = rng.standard_normal((3, 3)); ug -= np.eye(3) * np.trace(ug) / 3
    x = np.zeros(n, complex); x[0] = 1 / np.sqrt(4 * np.pi)
    for _ in range(50):
        x = o.step_euler(x, 0.02, ug)
    idx = {(int(l), int(m)): j for j, (l, m) in enumerate(zip(lm[0], lm[1]))}
    err = max(abs(x[idx[(l, -m)]] - (-1) ** abs(m) * np.conj(x[idx[(l, m)]])) for (l, m) in idx)
    assert err < 1e-15 and abs(np.trace(o.a2(x)) - 1) < 1e-14


def test_regularisation_and_cdrx_diagonals():
    o.init(12)
    D = np.diag([.5, .5, -1.])
    M = o.M_REG(D)
    assert np.count_nonzero(M - np.diag(np.diag(M))) == 0
    assert abs(M[-1, -1] + 10.6068117205577668 * np.sqrt(1.5)) < 1e-13       # l = L mode: -nu*||D||
    assert np.array_equal(np.diag(o.M_CDRX())[:6], [0, -6, -6, -6, -6, -6])


def test_apply_bounds():
    o.init(8)
    x = np.zeros(45, complex); x[0] = 1 / np.sqrt(4 * np.pi)
    x[3] = 5.0
    y = o.apply_bounds(x)
    assert abs(o.Sl(y, 2) / x[0].real ** 2 - 1) < 1e-14 and np.array_equal(y[6:], x[6:])


@pytest.mark.parametrize("tag,fn", [("v2", "a2_orth"), ("v4", "a4_orth"), ("c2b2", "a4_joint"), ("c2v2", "a4_jointcross")])
def test_orthotropic_moment_bodies(tag, fn):
    """src/moments.f90:242-311: the coefficient tensors extracted symbolically from the generated include bodies
    (tools/make_orthotropic_tables.py) against the numeric interpretation of the same text (tools/make_golden.py)"""
    for c in range(G["orth_b"].shape[0]):
        r = getattr(o, fn)(G["orth_b"][c], G["orth_n"][c])
        ref = G["orth_" + tag][c]
        assert r.shape == ref.shape
        assert np.abs(r - ref).max() < 5e-15 * np.abs(ref).max()


def test_orthotropic_enhancements_pins():
    """isotropic b, n, v distributions -> Eij = 1; index symmetrisers of src/tensorproducts.f90:53-82"""
    o.init(4)
    o._check_sym4()
    iso = np.zeros(15, complex); iso[0] = 1 / np.sqrt(4 * np.pi)
    e = np.eye(3)
    E = o.Eij_orthotropic(iso, iso, iso, e[0], e[1], e[2], (1, 1, 1, 1, 10, 1), 0.0, 1)
    assert np.abs(E - 1).max() < 1e-14
    # trace identities of the joint moments: <b^2 n^2 sin^2> contracts to 1, <v^2> has unit trace
    rng = np.random.default_rng(3)
    b = iso.copy(); n = iso.copy()
    b[1:6] = 0.02 * (rng.standard_normal(5)); n[1:6] = 0.02 * rng.standard_normal(5)
    for q in (b, n):        # impose n_l^-m = (-1)^m conj(n_l^m)
        q[1], q[2] = q[5].conjugate(), -q[4].conjugate(); q[3] = q[3].real
    assert abs(np.trace(o.a2_orth(b, n)) - 1) < 1e-6
    assert abs(np.einsum("iijj", o.a4_joint(b, n)) - 1) < 1e-6
    # n_grain /= 1 is "silently 0" in the reference's forward rheology -> 0/0
    assert np.all(np.isnan(o.Eij_orthotropic(iso, iso, iso, e[0], e[1], e[2], (1, 1, 1, 1, 10, 1), 0.0, 3)))


@pytest.mark.parametrize("tag,fn", [("c6", "a6"), ("c8", "a8")])
def test_high_order_structure_tensors(tag, fn):
    """src/moments.f90:220-236: coefficient tables of the unique a6 / a8 entries (tools/make_moment_tables.py) against
    the numeric interpretation of all 729 / 6561 assignments of the reference bodies (tools/make_golden.py)"""
    o.init(8)
    for c in range(G["hi_nlm"].shape[0]):
        r = getattr(o, fn)(G["hi_nlm"][c])
        ref = G["hi_" + tag][c]
        assert r.shape == ref.shape
        assert np.abs(r - ref).max() < 5e-15 * np.abs(ref).max()


def test_high_order_contractions_and_closures():
    """a8 -> a6 -> a4 -> a2 by contraction (float32-constant accuracy; the a4 alias quirk shows up in exactly one
    entry), delta function pins, isotropic n'=3 enhancements = 1"""
    o.init(8)
    from math import sqrt, pi
    d = np.zeros(45, complex)
    for l, j in ((0, 0), (2, 3), (4, 10), (6, 21), (8, 36)):
        d[j] = sqrt((2 * l + 1) / (4 * pi))                  # delta function along z
    assert abs(o.a8(d)[(2,) * 8] - 1) < 1e-7 and abs(o.a6(d)[(2,) * 6] - 1) < 1e-7
    assert abs(o.pfJ(d) - sum(2 * l + 1 for l in range(0, 9, 2))) < 1e-12
    x = np.zeros(45, complex); x[0] = 1 / sqrt(4 * pi)
    rng = np.random.default_rng(1)
    x[1:] = 0.02 * (rng.standard_normal(44) + 1j * rng.standard_normal(44))
    lm = [(l, m) for l in range(0, 9, 2) for m in range(-l, l + 1)]
    idx = {k: j for j, k in enumerate(lm)}
    for (l, m), j in idx.items():
        x[j] = x[j].real if m == 0 else ((-1) ** abs(m) * np.conj(x[idx[(l, -m)]]) if m < 0 else x[j])
    A8, A6, A4, A2 = o.a8(x), o.a6(x), o.a4(x), o.a2(x)
    assert np.abs(np.einsum("abcdefii", A8) - A6).max() < 1e-7
    dev = np.abs(np.einsum("abcdii", A6) - A4)
    assert np.argwhere(dev > 1e-7).tolist() == [[2, 1, 0, 1]]        # src/include/ev_c4__body.f90:78
    assert np.abs(np.einsum("abii", A4) - A2).max() < 1e-7
    iso = np.zeros(45, complex); iso[0] = 1 / sqrt(4 * pi)
    e = np.eye(3)
    for ng in (1, 3):
        assert np.abs(o.Eij_tranisotropic(iso, e[0], e[1], e[2], (1, 1e3), 0.0125, ng) - 1).max() < 1e-12
    # n' = -3 is not normalised by the reference: its isotropic denominator is I2*tau without the (1 + 2/15 cB + 2/3 cC) factor
    cA, cB, cC = o.rheo_params_tranisotropic((1, 1e3), 3, -3.0, 1)
    Es = 1 + 2 / 15 * cB + 2 / 3 * cC
    assert np.abs(o.Eij_tranisotropic(iso, e[0], e[1], e[2], (1, 1e3), 0.0, -3) - Es).max() < 1e-12 * abs(Es)
    S = np.diag([.5, .5, -1.])
    assert abs(o.ev_D4(iso, S) - 1) < 1e-6 and abs(o.E_CAFFE(iso, S, 0.1, 10, 1) - 1) < 1e-12
    assert abs(o.E_CAFFE(iso, S, 0.1, 10, 3) - 1) < 1e-6


@pytest.mark.parametrize("tag", ["a2", "a4", "a6"])
def test_state_ingest_maps(tag):
    """src/moments.f90:68-92: extracted affine maps against the numeric interpretation of the reference bodies, on
    arbitrary (non-symmetric) tensors so that every index the bodies read is pinned"""
    f = getattr(o, tag + "_to_nlm")
    for c in range(G["ingest_" + tag].shape[0]):
        r = f(G["ingest_" + tag][c])
        ref = G["ingest_" + tag + "_nlm"][c]
        assert np.abs(r - ref).max() < 5e-15 * np.abs(ref).max()


def test_state_ingest_round_trip():
    """SURVEY 8c pin (7): nlm - a4_to_nlm(a4(nlm)) is small (float32-constant accuracy), tests/ai-to-nlm/ai-to-nlm.py:95"""
    o.init(8)
    rng = np.random.default_rng(4)
    x = np.zeros(45, complex); x[0] = 1 / np.sqrt(4 * np.pi)
    lm = [(l, m) for l in range(0, 9, 2) for m in range(-l, l + 1)]
    idx = {k: j for j, k in enumerate(lm)}
    x[1:] = 0.03 * (rng.standard_normal(44) + 1j * rng.standard_normal(44))
    for (l, m), j in idx.items():
        x[j] = x[j].real if m == 0 else ((-1) ** abs(m) * np.conj(x[idx[(l, -m)]]) if m < 0 else x[j])
    assert np.abs(o.a2_to_nlm(o.a2(x)) - x[:6]).max() < 1e-7
    assert np.abs(o.a4_to_nlm(o.a4(x)) - x[:15]).max() < 1e-7
    assert np.abs(o.a6_to_nlm(o.a6(x)) - x[:28]).max() < 1e-7


def test_spectral_lrot_agrees_with_the_discrete_grain_ensemble():
    """SURVEY 8c pin (6): the reference's own cross-check -- a2 of the spectral M_LROT evolution against a2 of an ensemble
    of discrete grain axes rotated by ri_LROT (src/dynamics.f90:112-137) under the same flow (statistical agreement)"""
    L = 8
    lm, n = o.init(L)
    rng = np.random.default_rng(11)
    r0 = rng.standard_normal((4000, 3)); r0 /= np.linalg.norm(r0, axis=1)[:, None]     # isotropic ensemble
    ug = np.array([[0.4, 0.3, 0.0], [-0.1, 0.2, 0.0], [0.0, 0.1, -0.6]])            # compression + shear, traceless
    D, W = (ug + ug.T) / 2, (ug - ug.T) / 2
    Nt, dt = 41, 0.02
    ri = o.ri_LROT(r0, dt, Nt, np.tile(D, (Nt, 1, 1)), np.tile(W, (Nt, 1, 1)), 1.0)
    a2_disc = np.einsum("gi,gj->ij", ri[-1], ri[-1]) / ri.shape[1]
    a2_disc0 = np.einsum("gi,gj->ij", r0, r0) / r0.shape[0]
    x = np.zeros(n, complex); x[0] = 1 / np.sqrt(4 * np.pi)
    for _ in range(Nt - 1):
        x = o.step_euler(x, dt, ug, use_reg=False)
    a2_spec = o.a2(x)
    # remove the sampling noise of the initial ensemble (a2_disc0 - I/3) to first order
    assert np.abs(a2_disc - (a2_disc0 - np.eye(3) / 3) - a2_spec).max() < 0.02
    assert np.abs(a2_spec - np.eye(3) / 3).max() > 0.1          # the fabric did develop


# ---------------------------------------------------------------------------------------------------------------------
# Pinning against the COMPILED reference: tests/golden/ref_compiled.npz is produced by oracle/build_ref.sh +
# oracle/make_ref_fixtures.py on a machine with gfortran (this container has none, so the file may be absent).
# ---------------------------------------------------------------------------------------------------------------------
orc = o
REF_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_compiled.npz")


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_oracle_against_fixture(path, tol=1e-13):
    """Every hot-path procedure of the numpy oracle against stored outputs of specfabpy (or of the stand-in)."""
    d = np.load(path)
    GRAIN, ALPHA = (1.0, 1e3), 0.0125
    for L in (4, 8, 12, 20):
        orc.init(L)
        x, ug, tau = d["L%d_nlm" % L], d["L%d_ugrad" % L], d["L%d_tau" % L]
        D = (ug + ug.transpose(0, 2, 1)) / 2
        W = (ug - ug.transpose(0, 2, 1)) / 2
        assert np.array_equal(np.asarray(d["L%d_lm" % L]), np.array(orc.lm_list(L)).T)
        for p in range(4):
            assert _rel(orc.M_LROT(D[p], W[p], 1.0, 0.0), d["L%d_M_LROT" % L][p]) < tol
            assert _rel(orc.M_LROT(D[p], W[p], 0.7, 0.3), d["L%d_M_LROT_zeta" % L][p]) < tol
            assert _rel(orc.M_DDRX_src(tau[p]), d["L%d_M_DDRX_src" % L][p]) < tol          # qt**(2.0): complex pow in the reference
            assert _rel(orc.M_DDRX(x[p], tau[p]), d["L%d_M_DDRX" % L][p]) < tol
            assert _rel(orc.M_REG(D[p]), d["L%d_M_REG" % L][p]) < tol
            assert _rel(orc.a2(x[p]), d["L%d_a2" % L][p]) < tol
            assert _rel(orc.a4(x[p]), d["L%d_a4" % L][p]) < tol
        assert _rel(orc.M_CDRX(), d["L%d_M_CDRX" % L]) < tol
        e = np.eye(3)
        v, w = np.array([1.0, 2.0, -0.5]) / np.linalg.norm([1.0, 2.0, -0.5]), np.array([2.0, -1.0, 0.0]) / np.sqrt(5.0)
        tvw = np.outer(v, w) + np.outer(w, v)
        for p in range(3):
            assert _rel(orc.apply_bounds(4 * x[p]), d["L%d_apply_bounds" % L][p]) < tol
            ei, lami = orc.eig(x[p])
            assert _rel(lami, d["L%d_eig_lami" % L][p]) < 1e-12
            rei = d["L%d_eig_ei" % L][p]
            if np.min(np.abs(np.diff(np.sort(lami)))) > 1e-6:       # vectors only through their projectors, only when separated
                for i in range(3):
                    assert np.abs(np.outer(ei[i], ei[i]) - np.outer(rei[i], rei[i])).max() < 1e-9
            assert _rel(orc.Eij_tranisotropic(x[p], e[0], e[1], e[2], GRAIN, ALPHA, 1), d["L%d_Eij" % L][p]) < 1e-11
            assert _rel(orc.Eij_tranisotropic(x[p], rei[0], rei[1], rei[2], GRAIN, ALPHA, 1), d["L%d_Eij_eigframe" % L][p]) < 1e-11
            assert _rel(orc.Evw_tranisotropic(v, w, tvw, x[p], GRAIN, ALPHA, 1), d["L%d_Evw" % L][p]) < 1e-11
        # nlm_LROT: row t is the state before step t (src/dynamics.f90:99-110)
        n = x.shape[1]
        v0 = np.zeros(n, dtype=np.complex128)
        v0[0] = 1 / np.sqrt(4 * np.pi)
        traj = d["L%d_nlm_LROT" % L]
        for t in range(traj.shape[0]):
            assert _rel(v0, traj[t]) < 1e-12
            v0 = orc.step_euler(v0, 0.05, ug[0], use_reg=False)
    # the failed-dposv fallback branch (src/homogenizations.f90:174-185)
    orc.init(8)
    e = np.eye(3)
    xs, Es = d["fallback_nlm"], d["fallback_Eij"]
    nfb = 0
    for p in range(xs.shape[0]):
        if not np.isfinite(Es[p]).all():
            continue
        got, st = orc.Eij_tranisotropic(xs[p], e[0], e[1], e[2], GRAIN, ALPHA, 1, return_status=True)
        nfb += st == 1
        assert _rel(got, Es[p]) < (1e-6 if st == 1 else 1e-11)     # the regularised normal equations square the condition number
    return nfb


def test_ref_compiled_fixture():
    """Active as soon as somebody has run oracle/build_ref.sh + oracle/make_ref_fixtures.py on a gfortran host."""
    if not os.path.exists(REF_FIXTURE):
        pytest.skip("tests/golden/ref_compiled.npz absent: no Fortran compiler here; recipe: oracle/build_ref.sh, oracle/make_ref_fixtures.py")
    assert "compiled reference" in str(np.load(REF_FIXTURE)["source"])
    check_oracle_against_fixture(REF_FIXTURE)


def test_ref_fixture_pipeline(tmp_path):
    """The generator and this consumer work end to end (oracle stand-in behind specfabpy's signatures, temp file)."""
    import subprocess, sys
    out = str(tmp_path / "standin.npz")
    gen = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "make_ref_fixtures.py")
    subprocess.check_call([sys.executable, gen, "--standin", "--out", out])
    assert check_oracle_against_fixture(out) >= 0
    # the stand-in may never land on the golden path
    rc = subprocess.call([sys.executable, gen, "--standin", "--out", REF_FIXTURE], stderr=subprocess.DEVNULL)
    assert rc != 0 and (not os.path.exists(REF_FIXTURE) or "compiled reference" in str(np.load(REF_FIXTURE)["source"]))
