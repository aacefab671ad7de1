"""CPU: host-side logic of the kernel generator.  The block plans the CUDA code is emitted from are
interpreted here in numpy (same canonical/mirror bookkeeping, same forcing layout as the kernel's
prep_node) and compared with the oracle's dense operators."""
import numpy as np
import pytest

import specfab_oracle as o
from specfab_b200.codegen import emit_step as es
from specfab_b200.codegen.operators import Operators, CAT, idx
from util import random_states, random_ugrad, random_tau


def lane_forcing(qe, qo, g, sign):
    """forcing block of lane set A (sign=+1) / B (sign=-1), as prep_node() in sfb_step_kernel.cuh builds it"""
    f = np.zeros(23, complex)
    for D in range(-2, 3):
        f[es.f_qe(D)] = qe[sign * D + 2]
    for D in range(-1, 2):
        f[es.f_w(D)] = sign * 1j * qo[sign * D + 1]
    for k, (lk, mk) in enumerate(CAT):
        f[8 + k] = g[CAT.index((lk, sign * mk))]
    return f


def apply_plans(L, plans, y, qe, qo, g):
    """k = (M_LROT + M_DDRX_src(g)) y evaluated the way the generated code does"""
    n = (L + 1) * (L + 2) // 2
    out = np.zeros(n, complex)
    for sign in (+1, -1):
        f = lane_forcing(qe, qo, g, sign)
        for p in plans:
            acc = {li: 0j for li in p.rows}
            for (D, nu, used, items) in p.blocks:
                yv = {lj: y[idx(lj, sign * nu)] for lj in used}
                for (fidx, ent) in items:
                    for li, lst in ent.items():
                        acc[li] += f[fidx] * sum(c * yv[lj] for lj, c in lst)
            for li in p.rows:
                out[idx(li, sign * p.mu)] = acc[li]
    return out


@pytest.mark.parametrize("L", [4, 8, 12, 20])
@pytest.mark.parametrize("ddrx", [0, 1])
def test_block_plans_reproduce_the_dense_operators(L, ddrx):
    o.init(L)
    op, plans = es.plan(L, ddrx)
    y = random_states(L, 1, 3, False)[0]
    ug = random_ugrad(1, 4)[0]
    tau = random_tau(1, 5)[0]
    D, W = (ug + ug.T) / 2, (ug - ug.T) / 2
    epssq = D @ D
    qe = o.quad_rr(0.7 * D + (0.3 / np.sqrt(np.trace(epssq))) * epssq)
    qo = o.quad_tp(W)
    g = o.ddrx_weights(tau) if ddrx else np.zeros(15, complex)
    ref = o.M_LROT(D, W, 0.7, 0.3) @ y
    if ddrx:
        ref = ref + o.M_DDRX_src(tau) @ y
    got = apply_plans(L, plans, y, qe, qo, g)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 5e-15


def test_role_partition_covers_every_mu_once_and_is_balanced():
    for L, R in ((8, 2), (12, 2), (20, 4)):
        op, plans = es.plan(L, 1)
        roles, load = es.assign_roles(plans, R)
        mus = sorted(p.mu for r in roles for p in r)
        assert mus == list(range(L + 1))
        assert max(load) / (sum(load) / R) < 1.15


def test_physical_row_layout_is_a_bijection():
    for L in (4, 8, 20):
        rows = {es.phys_row(l, m) for l in range(0, L + 1, 2) for m in range(-l, l + 1)}
        assert len(rows) == (L + 1) * (L + 2) // 2 and max(rows) < es.nrow_phys(L)


def test_emitted_source_shape():
    body, tab, meta = es.emit(8, 1, 1, 64, "imm", True)
    assert body.count("SFB_ROW_OUT(") == 25          # canonical rows (l, mu>=0) at L=8
    assert body.count("SFB_LOCKSTEP(") == 9
    assert meta["dfma_node"] == 2 * sum(meta["dfma_role"])
    body2, tab2, meta2 = es.emit(8, 1, 2, 32, "cbank", False)
    assert "sfb_tab[" in body2 and tab2.startswith("__constant__ double sfb_tab[")


def test_operator_mirror_symmetries_hold_for_all_L():
    for L in (4, 6, 8, 10, 12, 14, 16, 18, 20):
        Operators(L)       # asserts the selection rules and exact mirror symmetries of the tables


def apply_plans_reduced(L, plans, y, qe, qo, g):
    """rows m >= 0 of (M_LROT + M_DDRX_src) y from the m >= 0 half of a real-ODF state only, the way the reduced
    kernels do it (emit_mu(reduced=True), sfb_step_loop_r.cuh): a column block with nu < 0 reads the rows (l_j, |nu|)
    and its partial sum is conj-mirrored, S = (-1)^nu conj(S')."""
    f = lane_forcing(qe, qo, g, +1)
    out = {}
    for p in plans:
        acc = {li: 0j for li in p.rows}
        for (D, nu, used, items) in p.blocks:
            yv = {lj: y[idx(lj, abs(nu))] for lj in used}          # positive plane only
            for (fidx, ent) in items:
                for li, lst in ent.items():
                    S = sum(c * yv[lj] for lj, c in lst)
                    if nu < 0:
                        S = (-1) ** abs(nu) * np.conj(S)
                    acc[li] += f[fidx] * S
        for li in p.rows:
            out[(li, p.mu)] = acc[li]
    return out


@pytest.mark.parametrize("L", [4, 8, 12])
@pytest.mark.parametrize("ddrx", [0, 1])
def test_reduced_form_plan_on_real_odf_states(L, ddrx):
    """the reduced kernels' algebra: M maps real-ODF states onto real-ODF states and its m >= 0 rows follow from the
    m >= 0 half of the state (csrc/sfb_step_kernel_r.cuh)"""
    o.init(L)
    op, plans = es.plan(L, ddrx)
    y = random_states(L, 1, 7, True)[0]
    ug, tau = random_ugrad(1, 8)[0], random_tau(1, 9)[0]
    D, W = (ug + ug.T) / 2, (ug - ug.T) / 2
    qe, qo = o.quad_rr(D), o.quad_tp(W)
    g = o.ddrx_weights(tau) if ddrx else np.zeros(15, complex)
    ref = o.M_LROT(D, W, 1.0, 0.0) @ y + (o.M_DDRX_src(tau) @ y if ddrx else 0)
    got = apply_plans_reduced(L, plans, y, qe, qo, g)
    scale = np.abs(ref).max()
    for (l, m), v in got.items():
        assert abs(v - ref[idx(l, m)]) < 5e-15 * scale
        assert abs((-1) ** m * np.conj(v) - ref[idx(l, -m)]) < 5e-15 * scale       # the mirror rows the kernel writes
        if m == 0:
            assert abs(v.imag) < 5e-15 * scale


@pytest.mark.parametrize("ddrx", [0, 1])
def test_inplace_stage_update_never_overwrites_a_live_row(ddrx):
    """in-place RK stages (emit(..., inplace=True)): replay the emitted order of column reads and commits"""
    L = 8
    body, _, _ = es.emit(L, ddrx, 1, 32, reduced=True, inplace=True)
    import re
    committed = set()
    rows_read_after_commit = []
    mu = None
    for line in body.splitlines():
        m = re.search(r"canonical mu = (\d+)", line)
        if m:
            mu = int(m.group(1))
        m = re.search(r"SFB_RROW_COMMIT\((\d+), (\d+),", line)
        if m:
            committed.add(es.pslot(int(m.group(1)), int(m.group(2))))
        m = re.search(r"= yp\[(\d+) \* SFB_TNR\]", line)
        if m and int(m.group(1)) in committed:
            rows_read_after_commit.append((mu, int(m.group(1))))
    assert not rows_read_after_commit
    assert len(committed) == (L // 2 + 1) ** 2          # every row is committed exactly once per stage
    assert body.count("SFB_RROW_COMMIT(") == len(committed)
