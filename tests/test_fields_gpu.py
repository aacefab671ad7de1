"""GPU parity of a2 / a4 / eigenframe / Eij against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

import specfab_oracle as orc
from util import random_states, random_ugrad, random_tau, relerr_nodes

pytestmark = pytest.mark.gpu
L = 8
GRAIN = (1.0, 1e3)       # ice 'linear': (Emm, Emt), alpha = 0.0125   src/specfabpy/constants.py:10
ALPHA = 0.0125


def evolved_states(N, seed, nsteps=120):
    """physically realisable states: evolve isotropy under random flow with the GPU step"""
    import specfab_b200 as sf
    lm, n = sf.init(L)
    x = np.zeros((N, n), dtype=np.complex128)
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    ug = random_ugrad(N, seed)
    tau = random_tau(N, seed + 1)
    rng = np.random.default_rng(seed + 2)
    nst = rng.integers(1, nsteps, N)
    out = x.copy()
    for k in sorted(set(nst.tolist())):
        pass
    # different strain per node: integrate in a few batches
    done = np.zeros(N, dtype=int)
    cur = x
    for target in (10, 40, nsteps):
        cur = sf.step_arr(cur, ug, tau, dt=0.01, Gamma0=1.0, terms=("lrot", "ddrx", "reg"), nsteps=target - done[0])
        done[:] = target
        sel = (nst <= target) & (nst > (0 if target == 10 else (10 if target == 40 else 40)))
        out[sel] = cur[sel]
    return out


def test_a2_a4_parity_and_quirk():
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    for physical in (True, False):
        x = random_states(L, 50, 11, physical)
        a2 = sf.a2_arr(x)
        a4 = sf.a4_arr(x)
        r2 = np.array([orc.a2(v) for v in x])
        r4 = np.array([orc.a4(v) for v in x])
        assert relerr_nodes(a2, r2).max() < 1e-13
        assert relerr_nodes(a4, r4).max() < 1e-13
        assert np.allclose(np.trace(a2, axis1=1, axis2=2), 1.0, atol=1e-13)
    # the reference's alias quirk ev(3,2,1,2)=ev(1,2,3,3) is reproduced (src/include/ev_c4__body.f90:78)
    assert np.array_equal(a4[:, 2, 1, 0, 1], a4[:, 0, 1, 2, 2])
    assert not np.allclose(a4[:, 2, 1, 0, 1], a4[:, 0, 1, 1, 2])
    # scalar API
    assert np.array_equal(sf.a2(x[3]), a2[3]) and np.array_equal(sf.a4(x[3]), a4[3])


def test_isotropic_pins():
    """SURVEY 8c pin (1): isotropic state -> a2 = I/3 exactly, Eij = 1"""
    import specfab_b200 as sf
    lm, n = sf.init(L)
    x = np.zeros((4, n), dtype=np.complex128)
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    assert np.array_equal(sf.a2_arr(x)[0], np.eye(3) / 3)
    e = np.tile(np.eye(3)[None], (4, 1, 1))
    E = sf.Eij_tranisotropic_arr(x, e[:, 0], e[:, 1], e[:, 2], GRAIN, ALPHA, 1)
    assert np.abs(E - 1).max() < 1e-12


def test_single_maximum_pin():
    """SURVEY 8c pin (3): delta function along z with ice 'linear' parameters -> E_mt ~ 9.97, E_mm ~ 0.00997"""
    import specfab_b200 as sf
    lm, n = sf.init(L)
    d = np.zeros((1, n), dtype=np.complex128)
    d[0, 0] = 1 / np.sqrt(4 * np.pi); d[0, 3] = np.sqrt(5 / (4 * np.pi)); d[0, 10] = 3 / np.sqrt(4 * np.pi)
    e = np.eye(3)
    E = sf.Eij_tranisotropic(d[0], e[0], e[1], e[2], GRAIN, ALPHA, 1)
    orc.init(L)
    ref = orc.Eij_tranisotropic(d[0], e[0], e[1], e[2], GRAIN, ALPHA, 1)
    assert np.abs(E / ref - 1).max() < 1e-9
    assert abs(E[3] - 9.97005242) < 1e-7 and abs(E[0] - 9.97005242e-3) < 1e-10


def test_eig_parity():
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    x = evolved_states(60, 21)
    ei, lami = sf.eig_arr(x)
    for p in range(x.shape[0]):
        e_ref, l_ref = orc.eig(x[p])
        assert np.abs(lami[p] - l_ref).max() < 1e-13
        gaps = min(abs(l_ref[0] - l_ref[1]), abs(l_ref[1] - l_ref[2]))
        if gaps > 1e-6:
            for i in range(3):
                assert np.abs(np.outer(ei[p, i], ei[p, i]) - np.outer(e_ref[i], e_ref[i])).max() < 1e-9 / gaps * 1e-4 + 1e-11
        assert np.abs(ei[p] @ ei[p].T - np.eye(3)).max() < 1e-14
    e1, l1 = sf.eig(x[5])
    assert np.array_equal(e1, ei[5]) and np.array_equal(l1, lami[5])


@pytest.mark.parametrize("plane", ["ij", "xy", "xz"])
def test_eigframe_planes(plane):
    import specfab_b200 as sf
    sf.init(L)
    rng = np.random.default_rng(5)
    A = rng.standard_normal((40, 3, 3))
    M = A + A.transpose(0, 2, 1)
    ei, lami = sf.eigframe_arr(M, plane)
    for p in range(M.shape[0]):
        e_ref, l_ref = orc.eigframe(M[p], plane)
        assert np.abs(lami[p] - l_ref).max() < 1e-12 * np.abs(l_ref).max()
        for i in range(3):
            assert np.abs(np.outer(ei[p, i], ei[p, i]) - np.outer(e_ref[i], e_ref[i])).max() < 1e-9
    with pytest.raises(sf.SpecfabB200Error):
        sf.eigframe_arr(M, "yz")


def test_Eij_parity_given_frames():
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    x = evolved_states(80, 31)
    rng = np.random.default_rng(32)
    # arbitrary orthonormal frames (not the eigenframe): Eij is defined for any (e1,e2,e3)
    Q = np.linalg.qr(rng.standard_normal((x.shape[0], 3, 3)))[0]
    for alpha in (ALPHA, 0.455, 0.0, 1.0):
        E, st = sf.Eij_tranisotropic_arr(x, Q[:, 0], Q[:, 1], Q[:, 2], GRAIN, alpha, 1, return_status=True)
        ref = np.array([orc.Eij_tranisotropic(x[p], Q[p, 0], Q[p, 1], Q[p, 2], GRAIN, alpha, 1) for p in range(x.shape[0])])
        assert (st == 0).all()
        assert np.abs(E / ref - 1).max() < 1e-9


def test_Eij_eigenframe_fused():
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    x = evolved_states(64, 41)
    E, ei, lami, st = sf.Eij_eigenframe_arr(x, GRAIN, ALPHA, 1, return_frame=True, return_status=True)
    # frames are only defined up to sign / degenerate rotation: feed the GPU's frame to the oracle
    ref = np.array([orc.Eij_tranisotropic(x[p], ei[p, 0], ei[p, 1], ei[p, 2], GRAIN, ALPHA, 1) for p in range(x.shape[0])])
    assert np.abs(E / ref - 1).max() < 1e-9
    ei2, lami2 = sf.eig_arr(x)
    assert np.array_equal(ei, ei2) and np.array_equal(lami, lami2)


def test_taylor_fallback_and_status():
    """Unphysical states make P indefinite: the reference falls back to the Tikhonov-regularised normal
    equations (src/homogenizations.f90:177-185).  Same branch, same numbers, plus a status flag."""
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    x = random_states(L, 40, 51, True, decay=1.0) * 4
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    e = np.tile(np.eye(3)[None], (x.shape[0], 1, 1))
    E, st = sf.Eij_tranisotropic_arr(x, e[:, 0], e[:, 1], e[:, 2], GRAIN, ALPHA, 1, return_status=True)
    nfb = 0
    conds = []
    for p in range(x.shape[0]):
        ref, rst = orc.Eij_tranisotropic(x[p], e[p, 0], e[p, 1], e[p, 2], GRAIN, ALPHA, 1, return_status=True)
        assert (st[p] & 3) == (1 if rst == 1 else (3 if rst == 2 else 0))
        if rst == 0:
            assert np.abs(E[p] - ref).max() <= 1e-9 * np.abs(ref).max()
        elif rst == 1:
            # The fallback solves (F^T F + 1e-6 I) x = F^T b with the partially factorised F: forming the normal equations
            # squares the condition number, so two correct FP64 evaluations (LAPACK's blocked dposv here, the kernel's unrolled
            # Cholesky there) legitimately differ by ~ eps * cond(F^T F + 1e-6 I).  The bound below is that figure with a
            # factor 100 of slack -- measured conditions are 1e5..1e8, i.e. 1e-9..1e-6 relative -- instead of a blanket 1e-7.
            P = orc.taylor_P(x[p], GRAIN, 1)
            c, _, info = orc._dposv(np.array(P, order="F"), np.zeros((6, 1), order="F"), lower=1)
            Pf = np.array(c)
            cond = np.linalg.cond(Pf.T @ Pf + orc.TIKHONOV_F * np.eye(6))
            conds.append(cond)
            assert info != 0 and np.abs(E[p] - ref).max() <= 100 * 2.2e-16 * cond * np.abs(ref).max(), (cond, np.abs(E[p] / ref - 1).max())
        nfb += rst == 1
    assert nfb > 0, "test did not exercise the fallback branch"
    with pytest.raises(sf.SpecfabB200Error):
        sf.Eij_tranisotropic_arr(x, e[:, 0], e[:, 1], e[:, 2], GRAIN, ALPHA, 2)      # "unsupported n'" (src/homogenizations.f90:110-112)


def test_moments_after_1000_steps():
    """north_star: 1e-9 on nlm, a2/a4 and Eij after 1000 FP64 steps"""
    import specfab_b200 as sf
    lm, n = sf.init(L)
    orc.init(L)
    N = 3
    x = np.zeros((N, n), dtype=np.complex128)
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    ug = random_ugrad(N, 61)
    ug[0] = np.diag([0.5, 0.5, -1.0])
    dt = -np.log(0.02) / 1000
    got = sf.step_arr(x, ug, dt=dt, terms=("lrot", "reg"), nsteps=1000)
    ref = x.copy()
    for p in range(N):
        v = ref[p]
        for _ in range(1000):
            v = orc.step_euler(v, dt, ug[p])
        ref[p] = v
    assert relerr_nodes(got, ref).max() < 1e-9
    assert relerr_nodes(sf.a2_arr(got), np.array([orc.a2(v) for v in ref])).max() < 1e-9
    assert relerr_nodes(sf.a4_arr(got), np.array([orc.a4(v) for v in ref])).max() < 1e-9
    E, ei, lami = sf.Eij_eigenframe_arr(got, GRAIN, ALPHA, 1, return_frame=True)
    Er = np.array([orc.Eij_tranisotropic(ref[p], ei[p, 0], ei[p, 1], ei[p, 2], GRAIN, ALPHA, 1) for p in range(N)])
    assert np.abs(E / Er - 1).max() < 1e-9


def test_apply_bounds_and_reduced_form():
    """src/dynamics.f90:530-557 and src/reducedform.f90:160-187"""
    import specfab_b200 as sf
    lm, n = sf.init(L)
    orc.init(L)
    x = random_states(L, 30, 71, True, decay=1.0) * 3       # strong fabrics: several exceed the delta-function spectrum
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    got = sf.apply_bounds_arr(x)
    ref = np.array([orc.apply_bounds(v) for v in x])
    assert np.abs(got - ref).max() < 1e-14 and (np.abs(ref - x).max(axis=1) > 1e-3).sum() > 3
    assert np.array_equal(sf.apply_bounds(x[2]), got[2])
    # reduced form: m >= 0 coefficients in (l, m) order; the round trip reproduces physical states exactly
    r = sf.nlm_to_rnlm_arr(x)
    assert r.shape == (30, sf.rnlm_len()) and sf.rnlm_len() == (L + 2) ** 2 // 4
    keep = [j for j in range(n) if lm[1, j] >= 0]
    assert np.array_equal(r, x[:, keep])
    back = sf.rnlm_to_nlm_arr(r)
    assert np.abs(back - x).max() < 1e-16
    assert np.array_equal(sf.rnlm_to_nlm(sf.nlm_to_rnlm(x[1])), back[1])
    # apply_bounds on the reduced-form field (what the FE coupler does per node): same factors as the full form, bit for bit
    import torch
    d = torch.from_numpy(np.ascontiguousarray(r.T)).cuda()
    o = sf.apply_bounds_rnlm_arr_dev(d, out=torch.empty_like(d))
    assert np.array_equal(o.cpu().numpy().T, sf.nlm_to_rnlm_arr(got))
    sf.apply_bounds_rnlm_arr_dev(d)                          # in place
    assert np.array_equal(d.cpu().numpy().T, sf.nlm_to_rnlm_arr(got))


# ---------------------------------------------------------------------------------------------
# Orthotropic grains (olivine): Eij_orthotropic_arr      src/specfabpy.f90:488-500
# ---------------------------------------------------------------------------------------------
OLIVINE = (1.0, 1.0, 1.0, 1.0, 1.0, 10.0)      # (Ebb, Enn, Evv, Env, Ebv, Enb): easy b-n slip system


def _frames(N, seed):
    rng = np.random.default_rng(seed)
    Q = np.linalg.qr(rng.standard_normal((N, 3, 3)))[0]
    return Q[:, :, 0].copy(), Q[:, :, 1].copy(), Q[:, :, 2].copy()


@pytest.mark.parametrize("third", ["derived", "given", "mixed", "null"])
def test_Eij_orthotropic_parity(third):
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    N = 37                      # not a multiple of the CTA's node tile
    q1 = random_states(L, N, 31, True, decay=0.35)
    q2 = random_states(L, N, 32, True, decay=0.35)
    q3 = random_states(L, N, 33, True, decay=0.35)
    if third == "derived":
        q3 = 0 * q3
    elif third == "mixed":
        q3[::3] = 0
    e1, e2, e3 = _frames(N, 34)
    E = sf.Eij_orthotropic_arr(q1, q2, None if third == "null" else q3, e1, e2, e3, OLIVINE, 0.0, 1)
    q3r = 0 * q1 if third == "null" else q3
    ref = np.array([orc.Eij_orthotropic(q1[p], q2[p], q3r[p], e1[p], e2[p], e3[p], OLIVINE, 0.0, 1) for p in range(N)])
    assert E.shape == (N, 6) and np.all(np.isfinite(E))
    assert np.abs(E / ref - 1).max() < 1e-11
    # scalar entry point
    assert np.array_equal(sf.Eij_orthotropic(q1[5], q2[5], q3r[5], e1[5], e2[5], e3[5], OLIVINE, 0.0, 1), E[5])
    if third == "null":      # nlm_3 = None is accepted by the scalar form too
        assert np.array_equal(sf.Eij_orthotropic(q1[5], q2[5], None, e1[5], e2[5], e3[5], OLIVINE, 0.0, 1), E[5])


def test_Eij_orthotropic_pins_and_errors():
    import specfab_b200 as sf
    lm, n = sf.init(L)
    x = np.zeros((9, n), dtype=np.complex128)
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    e1, e2, e3 = _frames(9, 2)
    E = sf.Eij_orthotropic_arr(x, x, x, e1, e2, e3, OLIVINE, 0.0, 1)
    assert np.abs(E - 1).max() < 1e-13
    # isotropic grains (all Eij_grain = 1) never enhance, whatever the fabric
    q1 = random_states(L, 9, 5, True, decay=0.35); q2 = random_states(L, 9, 6, True, decay=0.35)
    E = sf.Eij_orthotropic_arr(q1, q2, 0 * q1, e1, e2, e3, (1, 1, 1, 1, 1, 1), 0.0, 1)
    orc.init(L)
    ref = np.array([orc.Eij_orthotropic(q1[p], q2[p], 0 * q1[p], e1[p], e2[p], e3[p], (1, 1, 1, 1, 1, 1), 0.0, 1) for p in range(9)])
    assert np.abs(E / ref - 1).max() < 1e-11
    # n_grain /= 1: the reference's 0/0
    assert np.all(np.isnan(sf.Eij_orthotropic_arr(x, x, x, e1, e2, e3, OLIVINE, 0.0, 3)))
    with pytest.raises(ValueError):
        sf.Eij_orthotropic_arr(x, x, x, e1, e2, e3, (1.0, 10.0), 0.0, 1)
    assert sf.Eij_orthotropic_arr(x[:0], x[:0], x[:0], e1[:0], e2[:0], e3[:0], OLIVINE, 0.0, 1).shape == (0, 6)


def test_Eij_orthotropic_device_arrays():
    import torch
    import specfab_b200 as sf
    sf.init(L)
    N = 1000
    q1 = random_states(L, N, 41, True, decay=0.35); q2 = random_states(L, N, 42, True, decay=0.35)
    e1, e2, e3 = _frames(N, 43)
    host = sf.Eij_orthotropic_arr(q1, q2, None, e1, e2, e3, OLIVINE, 0.0, 1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    out = sf.Eij_orthotropic_arr_dev(dev(q1), dev(q2), None, dev(e1), dev(e2), dev(e3), OLIVINE, 0.0, 1)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().T, host)


# ---------------------------------------------------------------------------------------------
# a6 / a8 based fields: a6_arr, n'=3 Sachs, E_CAFFE, pfJ          
# ---------------------------------------------------------------------------------------------
def test_a6_parity():
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    x = random_states(L, 70, 51, True)
    a6 = sf.a6_arr(x)
    ref = np.array([orc.a6(v) for v in x])
    assert a6.shape == (70,) + (3,) * 6
    assert relerr_nodes(a6, ref).max() < 1e-13
    assert np.array_equal(sf.a6(x[3]), a6[3])
    assert np.array_equal(a6, a6.transpose(0, 2, 1, 3, 4, 6, 5))


@pytest.mark.parametrize("n_grain", [3, -3])
def test_Eij_tranisotropic_nonlinear_grains(n_grain):
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    N = 45
    x = evolved_states(N, 61)
    e1, e2, e3 = _frames(N, 62)
    for grain, alpha in ((GRAIN, ALPHA), ((0.5, 20.0), 0.3)):
        E, st = sf.Eij_tranisotropic_arr(x, e1, e2, e3, grain, alpha, n_grain, return_status=True)
        ref = np.array([orc.Eij_tranisotropic(x[p], e1[p], e2[p], e3[p], grain, alpha, n_grain) for p in range(N)])
        assert np.all(st == 0)
        assert np.abs(E / ref - 1).max() < 1e-10
    # fused eigenframe variant uses the same closure
    Ef, ei, lami = sf.Eij_eigenframe_arr(x, GRAIN, ALPHA, n_grain, return_frame=True)
    Eg = sf.Eij_tranisotropic_arr(x, ei[:, 0], ei[:, 1], ei[:, 2], GRAIN, ALPHA, n_grain)
    assert np.abs(Ef / Eg - 1).max() < 1e-12
    if n_grain == 3:          # isotropic pin (n' = -3 is not normalised by the reference)
        iso = np.zeros((3, x.shape[1]), dtype=np.complex128); iso[:, 0] = 1 / np.sqrt(4 * np.pi)
        assert np.abs(sf.Eij_tranisotropic_arr(iso, e1[:3], e2[:3], e3[:3], GRAIN, ALPHA, n_grain) - 1).max() < 1e-12


def test_Eij_n3_needs_L8():
    import specfab_b200 as sf
    lm, n = sf.init(6)
    x = np.zeros((2, n), dtype=np.complex128); x[:, 0] = 1 / np.sqrt(4 * np.pi)
    e = np.tile(np.eye(3)[None], (2, 1, 1))
    with pytest.raises((sf.SpecfabB200Error, ValueError)):
        sf.Eij_tranisotropic_arr(x, e[:, 0], e[:, 1], e[:, 2], GRAIN, ALPHA, 3)
    with pytest.raises(sf.SpecfabB200Error):
        sf.Eij_tranisotropic_arr(x, e[:, 0], e[:, 1], e[:, 2], GRAIN, ALPHA, 2)
    sf.init(L)


@pytest.mark.parametrize("n_grain", [1, 3])
def test_E_CAFFE_parity(n_grain):
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    N = 80
    x = evolved_states(N, 71)
    eps = random_tau(N, 72)
    E = sf.E_CAFFE_arr(x, eps, 0.1, 10.0, n_grain)
    ref = np.array([orc.E_CAFFE(x[p], eps[p], 0.1, 10.0, n_grain) for p in range(N)])
    ok = np.isfinite(ref)
    assert ok.sum() > N // 2 and np.array_equal(np.isfinite(E), ok)
    assert np.abs(E[ok] / ref[ok] - 1).max() < 1e-10
    assert (ref[ok] < 1).any() and (ref[ok] > 1).any()            # both branches of the Placidi law
    assert sf.E_CAFFE(x[2], eps[2], 0.1, 10.0, n_grain) == E[2]


def test_pfJ_parity():
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    x = random_states(L, 33, 81, False)
    for Lmax in (None, 4, 0):
        J = sf.pfJ_arr(x, Lmax)
        ref = np.array([orc.pfJ(v, Lmax) for v in x])
        assert np.abs(J / ref - 1).max() < 1e-14
    with pytest.raises((sf.SpecfabB200Error, ValueError)):
        sf.pfJ_arr(x, 10)


@pytest.mark.parametrize("rank", [2, 4, 6])
def test_state_ingest_parity(rank):
    """a2/a4/a6 -> nlm (src/moments.f90:68-92) on arbitrary tensors and as the inverse of a2/a4/a6"""
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    rng = np.random.default_rng(90 + rank)
    A = rng.standard_normal((50,) + (3,) * rank)
    got = getattr(sf, "a%d_to_nlm_arr" % rank)(A)
    ref = np.array([getattr(orc, "a%d_to_nlm" % rank)(a) for a in A])
    assert got.shape == ref.shape and relerr_nodes(got, ref).max() < 1e-14
    x = random_states(L, 20, 95, True, decay=0.4)
    back = getattr(sf, "a%d_to_nlm_arr" % rank)(getattr(sf, "a%d_arr" % rank)(x))
    assert np.abs(back - x[:, :back.shape[1]]).max() < 1e-7
    assert np.array_equal(getattr(sf, "a%d_to_nlm" % rank)(A[3]), got[3])


def test_step_moments_Eij_one_call():
    """step + a2 + a4 + eigenframe + Eij of the new state in one call == the separate calls"""
    import torch
    import specfab_b200 as sf
    sf.init(L)
    N = 300
    x = evolved_states(N, 91)
    ug, tau = random_ugrad(N, 92), random_tau(N, 93)
    kw = dict(dt=3.912e-3, Gamma0=4.0, terms=("lrot", "ddrx", "reg"), scheme="euler")
    d = sf.layout_nlm(torch.from_numpy(x).cuda())
    r = sf.step_moments_Eij_arr_dev(d, sf.layout_mat(torch.from_numpy(ug).cuda()), sf.layout_mat(torch.from_numpy(tau).cuda()),
                                    GRAIN, ALPHA, 1, out=torch.empty_like(d), want_a2=True, want_a4=True, want_frame=True, **kw)
    torch.cuda.synchronize()
    y = sf.step_arr(x, ug, tau, **kw)
    assert np.array_equal(r["nlm"].cpu().numpy().T, y)
    E, ei, lami = sf.Eij_eigenframe_arr(y, GRAIN, ALPHA, 1, return_frame=True)
    assert np.array_equal(r["Eij"].cpu().numpy().T, E)
    assert np.array_equal(r["lami"].cpu().numpy().T, lami)
    assert np.array_equal(r["a2"].cpu().numpy().transpose(2, 1, 0), sf.a2_arr(y))
    assert np.array_equal(r["a4"].cpu().numpy().transpose(4, 3, 2, 1, 0), sf.a4_arr(y))


@pytest.mark.parametrize("n_grain", [1, -3])
def test_step_moments_Eij_on_reduced_form_states(n_grain):
    """the FE time step on a field kept in reduced form (rows m >= 0): step_rnlm + a2 + eigenframe + Eij in one call ==
    the full-form calls bit for bit, and within tolerance of the oracle; the fields-only variant; n' = 3 is refused"""
    import torch
    import specfab_b200 as sf
    sf.init(L)
    N = 150
    x = evolved_states(N, 191)
    ug, tau = random_ugrad(N, 192), random_tau(N, 193)
    kw = dict(dt=3.912e-3, Gamma0=4.0, terms=("lrot", "ddrx", "reg"), scheme="rk4")
    rx = sf.nlm_to_rnlm_arr(x)
    d = torch.from_numpy(np.ascontiguousarray(rx.T)).cuda()
    ugd, taud = sf.layout_mat(torch.from_numpy(ug).cuda()), sf.layout_mat(torch.from_numpy(tau).cuda())
    r = sf.step_moments_Eij_rnlm_arr_dev(d, ugd, taud, GRAIN, ALPHA, n_grain, out=torch.empty_like(d), want_a2=True, want_frame=True, **kw)
    torch.cuda.synchronize()
    y = sf.step_arr(x, ug, tau, **kw)
    assert np.array_equal(r["rnlm"].cpu().numpy().T, sf.nlm_to_rnlm_arr(y))
    E, ei, lami = sf.Eij_eigenframe_arr(y, GRAIN, ALPHA, n_grain, return_frame=True)
    assert np.array_equal(r["Eij"].cpu().numpy().T, E)
    assert np.array_equal(r["lami"].cpu().numpy().T, lami)
    assert np.array_equal(r["ei"].cpu().numpy().transpose(2, 1, 0), ei)
    assert np.array_equal(r["a2"].cpu().numpy().transpose(2, 1, 0), sf.a2_arr(y))
    # oracle on a few nodes (full-form arithmetic of the reference)
    orc.init(L)
    for p in range(0, N, 37):
        ref = orc.step_rk4(x[p], 3.912e-3, ug[p], tau[p], Gamma0=4.0, use_ddrx=True)
        Er = orc.Eij_tranisotropic(ref, ei[p, 0], ei[p, 1], ei[p, 2], GRAIN, ALPHA, n_grain)
        assert np.abs(r["Eij"].cpu().numpy().T[p] / Er - 1).max() < 1e-9
    # fields only, on the stepped state
    f = sf.step_moments_Eij_rnlm_arr_dev(r["rnlm"], None, None, GRAIN, ALPHA, n_grain, want_a2=True, step=False)
    assert np.array_equal(f["Eij"].cpu().numpy(), r["Eij"].cpu().numpy()) and np.array_equal(f["a2"].cpu().numpy(), r["a2"].cpu().numpy())
    with pytest.raises(sf.SpecfabB200Error):
        sf.step_moments_Eij_rnlm_arr_dev(r["rnlm"], None, None, GRAIN, ALPHA, 3, step=False)


def test_Evw_tranisotropic_parity():
    """Evw_tranisotropic_arr (arbitrary v, w, tau per node) against the oracle's Evw_tranisotropic
    (src/enhancementfactors.f90:47-69); the six eigenframe pairs of Eij_tranisotropic are the special case."""
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    N = 41
    x = random_states(L, N, 81, True, decay=0.5)
    rng = np.random.default_rng(82)
    v = rng.standard_normal((N, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    w = rng.standard_normal((N, 3)); w /= np.linalg.norm(w, axis=1)[:, None]
    tau = random_tau(N, 83)
    for n_grain in (1, -3):
        got = sf.Evw_tranisotropic_arr(x, v, w, tau, GRAIN, ALPHA, n_grain)
        ref = np.array([orc.Evw_tranisotropic(v[p], w[p], tau[p], x[p], GRAIN, ALPHA, n_grain) for p in range(N)])
        assert np.abs(got / ref - 1).max() < 1e-9
    # Eij_tranisotropic = Evw with (e_i, e_j, tau_vv / tau_vw)   (src/enhancementfactors.f90:36-44, 398-413)
    e = np.tile(np.eye(3)[None], (N, 1, 1))
    E = sf.Eij_tranisotropic_arr(x, e[:, 0], e[:, 1], e[:, 2], GRAIN, ALPHA, 1)
    tvv = np.eye(3)[None] / 3 - np.einsum("pi,pj->pij", e[:, 2], e[:, 2])
    assert np.abs(sf.Evw_tranisotropic_arr(x, e[:, 2], e[:, 2], tvv, GRAIN, ALPHA, 1) / E[:, 2] - 1).max() < 1e-12
    tvw = np.einsum("pi,pj->pij", e[:, 0], e[:, 2]) + np.einsum("pi,pj->pij", e[:, 2], e[:, 0])
    assert np.abs(sf.Evw_tranisotropic_arr(x, e[:, 0], e[:, 2], tvw, GRAIN, ALPHA, 1) / E[:, 4] - 1).max() < 1e-12
    assert abs(sf.Evw_tranisotropic(x[0], v[0], w[0], tau[0], GRAIN, ALPHA, 1) - orc.Evw_tranisotropic(v[0], w[0], tau[0], x[0], GRAIN, ALPHA, 1)) < 1e-9 * abs(got[0]) + 1e-9
    with pytest.raises(sf.SpecfabB200Error):
        sf.Evw_tranisotropic_arr(x, v, w, tau, GRAIN, ALPHA, 3)
    with pytest.raises(ValueError):
        sf.Evw_tranisotropic_arr(x, v[:-1], w, tau, GRAIN, ALPHA, 1)
