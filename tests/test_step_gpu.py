"""GPU parity of the fused step kernels against the CPU oracle, through the C ABI.
Tolerances (BASELINE.json north_star): relative 1e-12 per step, 1e-9 after 1000 steps (FP64)."""
import numpy as np
import pytest

import specfab_oracle as orc
from util import random_states, random_ugrad, random_tau, relerr_nodes

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12
TOL_1000 = 1e-9


def oracle_steps(L, x, ug, tau, scheme, nsteps, **kw):
    orc.init(L)
    out = np.array(x, dtype=np.complex128)
    stepf = orc.step_rk4 if scheme == "rk4" else orc.step_euler
    dt = kw.pop("dt")
    for p in range(out.shape[0]):
        v = out[p]
        kwp = dict(kw)
        for key in ("Gamma0", "Lambda"):
            if np.ndim(kwp.get(key, 0.0)):
                kwp[key] = kwp[key][p]
        for _ in range(nsteps):
            v = stepf(v, dt, ug[p], None if tau is None else tau[p], **kwp)
        out[p] = v
    return out


def built_L():
    import specfab_b200 as sf
    return sorted({k["L"] for k in sf.build_info()["step_kernels"]})


@pytest.mark.parametrize("L", [4, 6, 8, 10, 12, 20])
@pytest.mark.parametrize("scheme", ["euler", "rk4"])
@pytest.mark.parametrize("physical", [True, False])
def test_step_lrot_reg(L, scheme, physical):
    import specfab_b200 as sf
    if L not in built_L():
        pytest.skip("L=%d not built" % L)
    N = 77 if L <= 12 else 21
    sf.init(L)
    x = random_states(L, N, 100 + L, physical)
    ug = random_ugrad(N, 200 + L)
    dt = 3.912e-3
    got = sf.step_arr(x, ug, dt=dt, iota=1.0, zeta=0.0, nu=1.0, terms=("lrot", "reg"), scheme=scheme)
    ref = oracle_steps(L, x, ug, None, scheme, 1, dt=dt, iota=1.0, zeta=0.0, nu_mult=1.0)
    assert relerr_nodes(got, ref).max() < TOL_STEP


@pytest.mark.parametrize("L", [4, 8, 10, 12, 20])
@pytest.mark.parametrize("scheme", ["euler", "rk4"])
@pytest.mark.parametrize("physical", [True, False])
def test_step_all_terms(L, scheme, physical):
    """physical states take the reduced (m >= 0) kernels, general complex states their in-kernel full fallback"""
    import specfab_b200 as sf
    if L not in built_L():
        pytest.skip("L=%d not built" % L)
    N = 45 if L <= 12 else 19
    sf.init(L)
    x = random_states(L, N, 300 + L, physical)
    ug = random_ugrad(N, 400 + L)
    tau = random_tau(N, 500 + L)
    dt = 3.912e-3
    kw = dict(iota=0.8, zeta=0.3, Gamma0=4.0, Lambda=0.1)
    got = sf.step_arr(x, ug, tau, dt=dt, nu=1.5, terms=("lrot", "ddrx", "cdrx", "reg"), scheme=scheme, **kw)
    ref = oracle_steps(L, x, ug, tau, scheme, 1, dt=dt, nu_mult=1.5, use_ddrx=True, use_cdrx=True, **kw)
    assert relerr_nodes(got, ref).max() < TOL_STEP


def test_step_ddrx_tau_defaults_to_D_and_pernode_rates():
    import specfab_b200 as sf
    L, N = 8, 33
    sf.init(L)
    x = random_states(L, N, 7, False)
    ug = random_ugrad(N, 8)
    rng = np.random.default_rng(9)
    G0 = rng.uniform(0.5, 5.0, N)
    Lam = rng.uniform(0.0, 0.5, N)
    D = (ug + ug.transpose(0, 2, 1)) / 2
    dt = 2e-3
    got = sf.step_arr(x, ug, None, dt=dt, Gamma0=G0, Lambda=Lam, terms=("lrot", "ddrx", "cdrx", "reg"))
    ref = oracle_steps(L, x, ug, D, "euler", 1, dt=dt, Gamma0=G0, Lambda=Lam, use_ddrx=True, use_cdrx=True)
    assert relerr_nodes(got, ref).max() < TOL_STEP


def test_terms_can_be_switched_off():
    import specfab_b200 as sf
    L, N = 8, 20
    sf.init(L)
    x = random_states(L, N, 17, True)
    ug = random_ugrad(N, 18)
    tau = random_tau(N, 19)
    dt = 1e-2
    got = sf.step_arr(x, ug, tau, dt=dt, Gamma0=3.0, terms=("ddrx",))
    ref = oracle_steps(L, x, ug, tau, "euler", 1, dt=dt, Gamma0=3.0, use_lrot=False, use_ddrx=True, use_reg=False)
    assert relerr_nodes(got, ref).max() < TOL_STEP
    got = sf.step_arr(x, ug, dt=dt, terms=("lrot",))
    ref = oracle_steps(L, x, ug, None, "euler", 1, dt=dt, use_reg=False)
    assert relerr_nodes(got, ref).max() < TOL_STEP


def test_config1_parcel_1000_euler_steps():
    """BASELINE config 1: L=8, LROT+REG, uniaxial compression, 1000 Euler steps from isotropy."""
    import specfab_b200 as sf
    L = 8
    lm, n = sf.init(L)
    N = 3
    x = np.zeros((N, n), dtype=np.complex128)
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    ug = np.zeros((N, 3, 3))
    ug[0] = np.diag([0.5, 0.5, -1.0])
    ug[1] = np.array([[0.0, 0.0, 1.0], [0, 0, 0], [0, 0, 0]])          # simple shear
    ug[2] = random_ugrad(1, 5)[0]
    dt = -np.log(0.02) / 1000
    got = sf.step_arr(x, ug, dt=dt, terms=("lrot", "reg"), nsteps=1000)
    ref = oracle_steps(L, x, ug, None, "euler", 1000, dt=dt)
    assert relerr_nodes(got, ref).max() < TOL_1000
    orc.init(L)
    a2 = orc.a2(got[0])
    assert abs(np.trace(a2) - 1) < 1e-12 and a2[2, 2] > 0.8     # single maximum along z
    # the same 1000 steps with the state kept in reduced form: identical rows, and still within tolerance of the oracle
    idx = [l * (l + 1) // 2 + m for l in range(0, L + 1, 2) for m in range(0, l + 1)]
    rgot = sf.step_rnlm_arr(sf.nlm_to_rnlm_arr(x), ug, dt=dt, terms=("lrot", "reg"), nsteps=1000)
    assert np.array_equal(rgot, got[:, idx])
    assert relerr_nodes(rgot, ref[:, idx]).max() < TOL_1000


def test_1000_steps_all_terms_rk4():
    import specfab_b200 as sf
    L, N = 8, 4
    sf.init(L)
    x = random_states(L, N, 31, True, decay=0.3)
    ug = random_ugrad(N, 32)
    tau = random_tau(N, 33)
    dt = 3.912e-3
    kw = dict(Gamma0=4.0, Lambda=0.05)
    got = sf.step_arr(x, ug, tau, dt=dt, terms=("lrot", "ddrx", "cdrx", "reg"), scheme="rk4", nsteps=250, **kw)
    ref = oracle_steps(L, x, ug, tau, "rk4", 250, dt=dt, use_ddrx=True, use_cdrx=True, **kw)
    assert relerr_nodes(got, ref).max() < TOL_1000
    idx = [l * (l + 1) // 2 + m for l in range(0, L + 1, 2) for m in range(0, l + 1)]
    rgot = sf.step_rnlm_arr(sf.nlm_to_rnlm_arr(x), ug, tau, dt=dt, terms=("lrot", "ddrx", "cdrx", "reg"), scheme="rk4", nsteps=250, **kw)
    assert np.array_equal(rgot, got[:, idx])


def test_empty_and_ragged_and_errors():
    import specfab_b200 as sf
    L = 8
    lm, n = sf.init(L)
    assert n == 45 and lm.shape == (2, 45) and tuple(lm[:, 1]) == (2, -2)
    out = sf.step_arr(np.zeros((0, n), complex), np.zeros((0, 3, 3)), dt=0.1)
    assert out.shape == (0, n)
    for N in (1, 15, 16, 17, 31, 32, 33, 100):
        x = random_states(L, N, N, True)
        ug = random_ugrad(N, N + 1)
        got = sf.step_arr(x, ug, dt=1e-3)
        ref = oracle_steps(L, x, ug, None, "euler", 1, dt=1e-3)
        assert relerr_nodes(got, ref).max() < TOL_STEP
    with pytest.raises(sf.SpecfabB200Error):
        sf.init(7)
    with pytest.raises(sf.SpecfabB200Error):
        sf.init(22)
    sf.init(L)
    with pytest.raises(ValueError):
        sf.step_arr(np.zeros((3, n - 1), complex), np.zeros((3, 3, 3)), dt=0.1)


def test_eps_zero_gives_nan_like_reference():
    """src/dynamics.f90:73: zeta/sqrt(tr eps^2) is 0/0 for eps == 0 -> NaN state (silent in the reference)."""
    import specfab_b200 as sf
    L = 8
    sf.init(L)
    x = random_states(L, 2, 3, True)
    ug = np.zeros((2, 3, 3))
    ug[1] = random_ugrad(1, 4)[0]
    got = sf.step_arr(x, ug, dt=1e-3)
    assert np.isnan(got[0]).any() and np.isfinite(got[1]).all()


def test_large_field_linearity_property():
    """Full-size property check (no oracle at this size): LROT+REG Euler is linear in nlm."""
    import specfab_b200 as sf
    L = 8
    sf.init(L)
    N = 200_000
    rng = np.random.default_rng(1)
    x1 = random_states(L, N, 41, True)
    x2 = random_states(L, N, 42, False)
    ug = random_ugrad(N, 43)
    a, b = 0.7, -1.3
    y1 = sf.step_arr(x1, ug, dt=5e-3, scheme="rk4")
    y2 = sf.step_arr(x2, ug, dt=5e-3, scheme="rk4")
    y12 = sf.step_arr(a * x1 + b * x2, ug, dt=5e-3, scheme="rk4")
    assert relerr_nodes(y12, a * y1 + b * y2).max() < 1e-12
    # spot-check 16 scattered nodes against the oracle
    sel = rng.choice(N, 16, replace=False)
    ref = oracle_steps(L, x1[sel], ug[sel], None, "rk4", 1, dt=5e-3)
    assert relerr_nodes(y1[sel], ref).max() < TOL_STEP


@pytest.mark.parametrize("L,terms", [(8, ("lrot", "reg")), (6, ("lrot", "ddrx", "reg")), (8, ("lrot", "ddrx", "reg")), (10, ("lrot", "reg"))])
@pytest.mark.parametrize("scheme", ["euler", "rk4"])
def test_reduced_kernel_and_its_fallback(L, terms, scheme):
    """The default kernels for these (L, terms) compute only the rows m >= 0 when a 32-node tile has the real-ODF
    symmetry bit for bit, and fall back to the full two-lane algorithm for the tile otherwise
    (csrc/sfb_step_kernel_r.cuh).  Mixed batches: symmetric tiles, general tiles, tiles with ONE asymmetric node,
    a ragged tail; in-place stepping; exact symmetry of the output."""
    import specfab_b200 as sf
    if L not in built_L():
        pytest.skip("L=%d not built" % L)
    lm, n = sf.init(L)
    T = 128                               # largest tile of the reduced kernels (32 one-lane, 64/128 two-lane)
    N = 4 * T + 13
    x = random_states(L, N, 700 + L, True)
    xg = random_states(L, N, 701 + L, False)
    x[T:T + 40] = xg[T:T + 40]            # general complex states
    x[T + 70] = xg[T + 70]                # one general node among symmetric ones
    x[2 * T + 5, 0] += 1e-3j              # Im n_0^0 != 0 only
    x[2 * T + 50, n - 1] += 1e-9          # one coefficient off its mirror by more than round-off
    x[2 * T + 100, 3] += 1e-9j
    x[3 * T + 40, 3] += 1e-18j            # round-off sized asymmetry (treated as a real ODF)
    ug = random_ugrad(N, 702 + L)
    tau = random_tau(N, 703 + L)
    dt = 3.912e-3
    ddrx = "ddrx" in terms
    got = sf.step_arr(x, ug, tau if ddrx else None, dt=dt, Gamma0=4.0, terms=terms, scheme=scheme)
    ref = oracle_steps(L, x, ug, tau if ddrx else None, scheme, 1, dt=dt, Gamma0=4.0, use_ddrx=ddrx)
    assert relerr_nodes(got, ref).max() < TOL_STEP
    # symmetric nodes stay symmetric bit for bit (so the next step takes the reduced path again)
    idx = {k: j for j, k in enumerate(zip(lm[0].tolist(), lm[1].tolist()))}
    sym_nodes = [p for p in range(N) if p < T or p >= 3 * T]         # tiles that certainly took the reduced path
    for (l, m), j in idx.items():
        if m > 0:
            assert np.array_equal(got[sym_nodes, idx[(l, -m)]], (-1) ** m * np.conj(got[sym_nodes, j]))
        if m == 0:
            assert np.all(got[sym_nodes, j].imag == 0)
    # in-place on the device gives the same bits
    import torch
    d = sf.layout_nlm(torch.from_numpy(x).cuda())
    dug = sf.layout_mat(torch.from_numpy(ug).cuda())
    dtau = sf.layout_mat(torch.from_numpy(tau).cuda()) if ddrx else None
    sf.step_arr_dev(d, dug, dtau, dt=dt, Gamma0=4.0, terms=terms, scheme=scheme, out=d)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().T, got)


@pytest.mark.parametrize("L,terms", [(8, ("lrot", "reg")), (8, ("lrot", "ddrx", "reg")), (12, ("lrot", "reg")), (12, ("lrot", "ddrx", "reg")),
                                     (20, ("lrot", "ddrx", "cdrx", "reg"))])
def test_full_form_kernels_agree_with_the_defaults(L, terms):
    """variant 40 = the stand-alone full-form kernel of every (L, terms) (build.py FULL_DEFAULT): same results as the
    default (reduced) kernels on physical states, to round-off, and the same oracle parity on general states"""
    import specfab_b200 as sf
    from specfab_b200 import _lib
    if L not in built_L():
        pytest.skip("L=%d not built" % L)
    have = {(k["L"], k["ddrx"], k["variant"]) for k in sf.build_info()["step_kernels"]}
    if (L, int("ddrx" in terms), 40) not in have:
        pytest.skip("no separate full-form kernel for this configuration")
    sf.init(L)
    N = 150 if L <= 12 else 40
    ug, tau = random_ugrad(N, 800 + L), random_tau(N, 801 + L)
    ddrx = "ddrx" in terms
    kw = dict(dt=3.912e-3, Gamma0=4.0, Lambda=0.1, terms=terms)
    try:
        for scheme in ("euler", "rk4"):
            xs = random_states(L, N, 802 + L, True)
            _lib.load().sfb_set_variant(0)
            a = sf.step_arr(xs, ug, tau, scheme=scheme, **kw)
            _lib.load().sfb_set_variant(40)
            b = sf.step_arr(xs, ug, tau, scheme=scheme, **kw)
            assert relerr_nodes(a, b).max() < 1e-14
        xg = random_states(L, 24, 803 + L, False)
        got = sf.step_arr(xg, ug[:24], tau[:24], scheme="rk4", **kw)
        ref = oracle_steps(L, xg, ug[:24], tau[:24], "rk4", 1, dt=3.912e-3, Gamma0=4.0, Lambda=0.1, use_ddrx=ddrx, use_cdrx="cdrx" in terms)
        assert relerr_nodes(got, ref).max() < TOL_STEP
    finally:
        _lib.load().sfb_set_variant(0)


# --------------------------------------------------------------------------------------------------
# reduced-form states (sfb_step_rnlm_arr): rows m >= 0 in, rows m >= 0 out  (src/reducedform.f90:160-187)
# --------------------------------------------------------------------------------------------------
def oracle_rnlm_index(L):
    """positions of the m >= 0 coefficients in nlm, in rnlm order (l ascending, m = 0..l)"""
    lm = [(l, m) for l in range(0, L + 1, 2) for m in range(-l, l + 1)]
    return [lm.index((l, m)) for l in range(0, L + 1, 2) for m in range(0, l + 1)]


@pytest.mark.parametrize("L,terms", [(4, ("lrot", "reg")), (4, ("lrot", "ddrx", "reg")), (6, ("lrot", "ddrx", "reg")), (8, ("lrot", "reg")),
                                     (8, ("lrot", "ddrx", "reg")), (10, ("lrot", "reg")), (12, ("lrot", "ddrx", "cdrx", "reg")),
                                     (16, ("lrot", "reg")), (20, ("lrot", "ddrx", "cdrx", "reg"))])
@pytest.mark.parametrize("scheme", ["euler", "rk4"])
def test_step_on_reduced_form_states(L, terms, scheme):
    """step_rnlm_arr == nlm_to_rnlm(step_arr(rnlm_to_nlm(rnlm))) bit for bit (the same arithmetic on the same rows), and
    within 1e-12 of the oracle's full-form step; ragged batch (tile edges), host and device entry points."""
    import torch
    import specfab_b200 as sf
    if L not in built_L():
        pytest.skip("L not built")
    lm, n = sf.init(L)
    N = 77
    x = random_states(L, N, 5 + L, True)
    ug, tau = random_ugrad(N, 6), random_tau(N, 7)
    kw = dict(dt=2e-3, Gamma0=3.0, Lambda=0.7, terms=terms, scheme=scheme)
    idx = oracle_rnlm_index(L)
    r = sf.rnlm_len()
    assert r == len(idx) == (L + 2) ** 2 // 4
    rx = sf.nlm_to_rnlm_arr(x)
    assert np.array_equal(rx, x[:, idx])
    full = sf.step_arr(x, ug, tau, **kw)
    got = sf.step_rnlm_arr(rx, ug, tau, **kw)
    assert got.shape == (N, r)
    assert np.array_equal(got, full[:, idx]), "reduced-form step differs from the m >= 0 rows of the full-form step"
    assert np.array_equal(sf.rnlm_to_nlm_arr(got), full)
    k = N if L <= 12 else 10           # the pure-Python oracle is slow at large L
    ref = oracle_steps(L, x[:k], ug[:k], tau[:k], scheme, 1, dt=2e-3, Gamma0=3.0, Lambda=0.7, use_ddrx="ddrx" in terms, use_cdrx="cdrx" in terms)
    assert relerr_nodes(got[:k], ref[:, idx]).max() < TOL_STEP
    # device-resident entry point, three sub-steps, in place
    d = torch.from_numpy(np.ascontiguousarray(rx.T)).cuda()
    sf.step_rnlm_arr_dev(d, sf.layout_mat(torch.from_numpy(ug).cuda()), sf.layout_mat(torch.from_numpy(tau).cuda()), nsteps=3, **kw)
    full3 = sf.step_arr(x, ug, tau, nsteps=3, **kw)
    assert np.array_equal(d.cpu().numpy().T, full3[:, idx])


def test_reduced_form_step_errors_and_empty():
    import specfab_b200 as sf
    L = 8
    lm, n = sf.init(L)
    r = sf.rnlm_len()
    assert sf.step_rnlm_arr(np.zeros((0, r), complex), np.zeros((0, 3, 3)), dt=0.1).shape == (0, r)
    with pytest.raises(ValueError):
        sf.step_rnlm_arr(np.zeros((3, n), complex), np.zeros((3, 3, 3)), dt=0.1)      # full-form array handed to the reduced entry
    from specfab_b200 import _lib
    _lib.load().sfb_set_variant(40)        # full-form kernels have no reduced I/O: must be refused, not silently wrong
    try:
        with pytest.raises(sf.SpecfabB200Error):
            sf.step_rnlm_arr(sf.nlm_to_rnlm_arr(random_states(L, 4, 1, True)), random_ugrad(4, 2), dt=1e-3)
    finally:
        _lib.load().sfb_set_variant(0)


def test_pin_array_in_place():
    """page-locking the caller's own arrays (sfb_host_register) changes the speed of the host-pointer path, not its result"""
    import specfab_b200 as sf
    L, N = 8, 5000
    lm, n = sf.init(L)
    x = np.asfortranarray(random_states(L, N, 41, True))
    ug = np.asfortranarray(random_ugrad(N, 42))
    ref = sf.step_arr(x, ug, dt=1e-3, scheme="rk4")
    out = np.empty((N, n), dtype=np.complex128, order="F")
    for a in (x, ug, out):
        sf.pin_array(a)
    try:
        got = sf.step_arr(x, ug, dt=1e-3, scheme="rk4", out=out)
        assert got is out and np.array_equal(out, ref)
    finally:
        for a in (x, ug, out):
            sf.unpin_array(a)
    with pytest.raises(ValueError):
        sf.pin_array(np.zeros((4, 4))[::2])


def test_multi_device_entry_point_and_general_flag():
    """sfb_step_arr_multi (contiguous node ranges on 32-node boundaries, one host thread per device): bit-identical to the
    one-device call for every device list the box offers; error paths; SFB_STEP_GENERAL."""
    import specfab_b200 as sf
    from specfab_b200 import _lib
    L = 8
    sf.init(L)
    nd = _lib.load().sfb_device_count()
    N = 5_003
    x = random_states(L, N, 71, True)
    ug, tau = random_ugrad(N, 72), random_tau(N, 73)
    g0 = np.linspace(1.0, 4.0, N)
    kw = dict(dt=3e-3, Gamma0=g0, terms=("lrot", "ddrx", "reg"), nsteps=3)
    ref = sf.step_arr(x, ug, tau, **kw)
    lists = [[0]] + ([[0, 1], [1, 0]] if nd >= 2 else []) + ([list(range(nd))] if nd > 2 else [])
    for devs in lists:
        got = sf.step_arr(x, ug, tau, devices=devs, **kw)
        assert np.array_equal(got, ref), devs
    idx = [l * (l + 1) // 2 + m for l in range(0, L + 1, 2) for m in range(0, l + 1)]
    rgot = sf.step_rnlm_arr(sf.nlm_to_rnlm_arr(x), ug, tau, devices=lists[-1], **kw)
    assert np.array_equal(rgot, ref[:, idx])
    for bad in ([], [nd], [0, 0], [-1]):
        with pytest.raises(sf.SpecfabB200Error):
            sf.step_arr(x, ug, tau, devices=bad, **kw)
    # the general path on physical states: same result to round-off, and independent of the batch composition
    kw1 = dict(dt=3e-3, Gamma0=2.0, terms=("lrot", "ddrx", "reg"))
    a = sf.step_arr(x, ug, tau, general=True, **kw1)
    b = sf.step_arr(x, ug, tau, **kw1)
    assert relerr_nodes(a, b).max() < 1e-13
    sel = np.arange(0, N, 7)
    assert np.array_equal(sf.step_arr(x[sel], ug[sel], tau[sel], general=True, **kw1), a[sel])
    with pytest.raises(sf.SpecfabB200Error):
        o = sf._opts(1e-3, 1.0, 0.0, 1.0, 0.0, 0.0, ("lrot",), "euler", 1)
        o.reserved = 2
        _lib.check(_lib.load().sfb_step_arr(x.ctypes.data, x.ctypes.data, 1, 1, ug.ctypes.data, None, __import__("ctypes").byref(o)))
