"""CPU: the dense C restatement (timed CPU baseline) agrees with the numpy oracle."""
import numpy as np
import pytest

import specfab_oracle as o
import oracle_c as oc
from util import random_states, random_ugrad, random_tau, relerr_nodes


@pytest.mark.parametrize("L", [4, 8, 12])
@pytest.mark.parametrize("scheme", ["euler", "rk4"])
def test_c_oracle_matches_numpy_oracle(L, scheme):
    o.init(L)
    oc.init(L)
    N = 10
    x = random_states(L, N, 1, False)
    ug = random_ugrad(N, 2)
    tau = random_tau(N, 3)
    kw = dict(iota=0.9, zeta=0.2, nu_mult=1.3, Gamma0=3.0, Lambda=0.2, use_ddrx=True, use_cdrx=True)
    stepf = o.step_rk4 if scheme == "rk4" else o.step_euler
    ref = np.array([stepf(x[p], 3e-3, ug[p], tau[p], **kw) for p in range(N)])
    got = oc.step_batch(x, ug, tau, dt=3e-3, scheme=scheme, **kw)
    assert relerr_nodes(got, ref).max() < 1e-14


def test_c_oracle_lrot_only_many_steps():
    L = 8
    o.init(L)
    oc.init(L)
    x = random_states(L, 3, 5, True)
    ug = random_ugrad(3, 6)
    ref = x.copy()
    for p in range(3):
        v = ref[p]
        for _ in range(40):
            v = o.step_euler(v, 5e-3, ug[p], use_reg=False)
        ref[p] = v
    got = oc.step_batch(x, ug, dt=5e-3, use_reg=False, nsteps=40)
    assert relerr_nodes(got, ref).max() < 1e-13
