"""GPU parity of the operator-export entry points (matrix form of M_LROT / M_DDRX / M_REG / M_CDRX)."""
import numpy as np
import pytest

import specfab_oracle as orc
from util import random_states, random_ugrad, random_tau

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [4, 8, 12, 20])
def test_operator_matrices_match_oracle(L):
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    N = 5
    ug = random_ugrad(N, 10 + L)
    tau = random_tau(N, 20 + L)
    x = random_states(L, N, 30 + L, True)
    D = (ug + ug.transpose(0, 2, 1)) / 2
    W = (ug - ug.transpose(0, 2, 1)) / 2
    M = sf.M_LROT_arr(D, W, 0.8, 0.25)
    Ms = sf.M_DDRX_src_arr(tau)
    Md = sf.M_DDRX_arr(x, tau)
    Mr = sf.M_REG_arr(D)
    for p in range(N):
        ref = orc.M_LROT(D[p], W[p], 0.8, 0.25)
        assert np.abs(M[p] - ref).max() < 1e-14 * np.abs(ref).max()
        ref = orc.M_DDRX_src(tau[p])
        assert np.abs(Ms[p] - ref).max() < 1e-14 * np.abs(ref).max()
        ref = orc.M_DDRX(x[p], tau[p])
        assert np.abs(Md[p] - ref).max() < 1e-13 * np.abs(ref).max()
        ref = orc.M_REG(D[p])
        assert np.abs(Mr[p] - ref).max() < 1e-14 * np.abs(ref).max()
    assert np.array_equal(sf.M_CDRX(x[0]).real, orc.M_CDRX())
    # scalar forms with the reference's signatures
    assert np.array_equal(sf.M_LROT(x[0], D[0], W[0], 0.8, 0.25), M[0])
    assert np.array_equal(sf.M_DDRX(x[1], tau[1]), Md[1])
    assert np.array_equal(sf.M_DDRX_src(x[1], tau[1]), Ms[1])
    assert np.array_equal(sf.M_REG(x[2], D[2]), Mr[2])


def test_matrix_form_agrees_with_fused_step():
    """nlm + dt*(M_LROT + Gamma0*M_DDRX + M_REG) @ nlm assembled from the exported matrices == step_arr"""
    import specfab_b200 as sf
    L, N = 8, 6
    sf.init(L)
    x = random_states(L, N, 1, False)
    ug = random_ugrad(N, 2)
    tau = random_tau(N, 3)
    D = (ug + ug.transpose(0, 2, 1)) / 2
    W = (ug - ug.transpose(0, 2, 1)) / 2
    M = sf.M_LROT_arr(D, W, 1.0, 0.0) + 4.0 * sf.M_DDRX_arr(x, tau) + sf.M_REG_arr(D)
    ref = x + 1e-2 * np.einsum("pij,pj->pi", M, x)
    got = sf.step_arr(x, ug, tau, dt=1e-2, Gamma0=4.0, terms=("lrot", "ddrx", "reg"))
    assert (np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)).max() < 1e-13


def test_nlm_LROT_integrator():
    """src/dynamics.f90:99-110"""
    import specfab_b200 as sf
    L = 8
    lm, n = sf.init(L)
    orc.init(L)
    Nt = 30
    ug = random_ugrad(Nt, 5)
    D = (ug + ug.transpose(0, 2, 1)) / 2
    W = (ug - ug.transpose(0, 2, 1)) / 2
    nlm0 = np.zeros(n, complex)
    nlm0[0] = 1 / np.sqrt(4 * np.pi)
    got = sf.nlm_LROT(nlm0, 0.02, Nt, D, W, 1.0)
    ref = orc.nlm_LROT(nlm0, 0.02, Nt, D, W, 1.0)
    assert np.abs(got - ref).max() < 1e-13


@pytest.mark.parametrize("L", [4, 8, 12])
def test_reduced_operators(L):
    """src/reducedform.f90:76-120: Mrr, Mri, Mir, Mii written directly, and reduce_M of a dense operator"""
    import specfab_b200 as sf
    sf.init(L)
    orc.init(L)
    N = 7
    ug = random_ugrad(N, 40 + L)
    tau = random_tau(N, 50 + L)
    x = random_states(L, N, 60 + L, True)
    D = (ug + ug.transpose(0, 2, 1)) / 2
    W = (ug - ug.transpose(0, 2, 1)) / 2
    r = sf.rnlm_len()
    cases = [(sf.M_LROT_reduced_arr(D, W, 0.8, 0.25), lambda p: orc.M_LROT(D[p], W[p], 0.8, 0.25)),
             (sf.M_DDRX_reduced_arr(x, tau), lambda p: orc.M_DDRX(x[p], tau[p])),
             (sf.M_DDRX_reduced_arr(None, tau, src_only=True), lambda p: orc.M_DDRX_src(tau[p])),
             (sf.reduce_M_arr(sf.M_REG_arr(D)), lambda p: orc.M_REG(D[p])),
             (sf.reduce_M_arr(sf.M_LROT_arr(D, W, 0.8, 0.25)), lambda p: orc.M_LROT(D[p], W[p], 0.8, 0.25))]
    for got, ref_of in cases:
        assert all(g.shape == (N, r, r) for g in got)
        for p in range(N):
            Mfull = ref_of(p)
            ref = orc.reduce_M(Mfull)
            scale = np.abs(Mfull).max()
            for g, rf in zip(got, ref):
                assert np.abs(g[p] - rf).max() < 1e-13 * scale
    # the reduced operators reproduce the full tendency on a physical state
    Mrr, Mri, Mir, Mii = sf.M_LROT_reduced_arr(D, W, 1.0, 0.0)
    rn = sf.nlm_to_rnlm_arr(x)
    drn = np.einsum("pij,pj->pi", Mrr, rn.real) + np.einsum("pij,pj->pi", Mri, rn.imag) \
        + 1j * (np.einsum("pij,pj->pi", Mir, rn.real) + np.einsum("pij,pj->pi", Mii, rn.imag))
    dn = np.einsum("pij,pj->pi", sf.M_LROT_arr(D, W, 1.0, 0.0), x)
    assert np.abs(drn - sf.nlm_to_rnlm_arr(dn)).max() < 1e-13 * np.abs(dn).max()
    # scalar form with the reference's signature
    one = sf.reduce_M(sf.M_LROT(x[0], D[0], W[0], 1.0, 0.0), r)
    assert all(np.array_equal(a, b[0]) for a, b in zip(one, (Mrr, Mri, Mir, Mii)))
