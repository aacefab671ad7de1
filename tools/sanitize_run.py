"""Small run of every default kernel family (reduced path and general-state fallback) for compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import specfab_b200 as sf
from util import random_states, random_ugrad, random_tau

N = 200
for L in (6, 8, 12, 20):
    sf.init(L)
    for physical in (True, False):
        x = random_states(L, N, 1 + L, physical)
        ug, tau = random_ugrad(N, 2), random_tau(N, 3)
        for terms in (("lrot", "reg"), ("lrot", "ddrx", "cdrx", "reg")):
            for scheme in ("euler", "rk4"):
                y = sf.step_arr(x, ug, tau, dt=1e-3, Gamma0=2.0, Lambda=0.1, terms=terms, scheme=scheme)
                assert np.all(np.isfinite(y))
    if L == 8:
        x = random_states(L, N, 5, True, 0.35)
        e = np.linalg.qr(np.random.default_rng(0).standard_normal((N, 3, 3)))[0]
        sf.Eij_eigenframe_arr(x, (1.0, 1e3), 0.0125, 1)
        sf.Eij_eigenframe_arr(x, (1.0, 1e3), 0.0125, 3)
        sf.Eij_orthotropic_arr(x, x, None, e[:, :, 0], e[:, :, 1], e[:, :, 2], (1, 1, 1, 1, 1, 10), 0.0, 1)
        sf.E_CAFFE_arr(x, tau, 0.1, 10.0, 3)
        sf.M_LROT_reduced_arr(tau, tau - tau.transpose(0, 2, 1), 1.0, 0.0)
        sf.a6_arr(x); sf.a4_to_nlm_arr(sf.a4_arr(x))
print("sanitize_run ok")
