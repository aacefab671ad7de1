"""Time the HBM-bound per-node field kernels on resident inputs and print their algorithmic bandwidth against the measured
HBM peak: a2, a4, a6, eig, nlm<->rnlm, apply_bounds, reduce_M, M_REG (L = 8)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import specfab_b200 as sf
from specfab_b200 import _lib
from util import random_states, random_tau

try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
lib = _lib.load()
L = 8
lm, n = sf.init(L)
r = sf.rnlm_len()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x = torch.from_numpy(np.ascontiguousarray(random_states(L, N, 1, True, 0.35).T)).cuda()
xr = torch.empty((r, N), dtype=torch.complex128, device="cuda")
xo = torch.empty_like(x)
eps = torch.from_numpy(np.ascontiguousarray(random_tau(N, 2).transpose(2, 1, 0))).cuda()
f64 = lambda *shape: torch.empty(shape, dtype=torch.float64, device="cuda")
a2, a4, ei, lami = f64(9, N), f64(81, N), f64(9, N), f64(3, N)
Nm = 50_000
Mc = torch.zeros((n, n, Nm), dtype=torch.complex128, device="cuda")
Mr = f64(n, n, Nm)
R4 = [f64(r, r, Nm) for _ in range(4)]
Na6 = 200_000
a6 = f64(729, Na6)
P = lambda t: t.data_ptr()
cases = [
    ("a2_arr", N, 6 * 16 + 72, lambda: lib.sfb_a2_arr_dev(P(x), N, N, P(a2), None)),
    ("a4_arr", N, 15 * 16 + 648, lambda: lib.sfb_a4_arr_dev(P(x), N, N, P(a4), None)),
    ("a6_arr", Na6, 28 * 16 + 729 * 8, lambda: lib.sfb_a6_arr_dev(P(x), Na6, N, P(a6), None)),
    ("eig_arr", N, 6 * 16 + 96, lambda: lib.sfb_eig_arr_dev(P(x), N, N, P(ei), P(lami), None)),
    ("nlm_to_rnlm", N, 16 * (r + r), lambda: lib.sfb_nlm_to_rnlm_arr_dev(P(x), P(xr), N, N, N, None)),
    ("rnlm_to_nlm", N, 16 * (r + n), lambda: lib.sfb_rnlm_to_nlm_arr_dev(P(xr), P(xo), N, N, N, None)),
    ("apply_bounds (out of place)", N, 16 * 2 * n, lambda: lib.sfb_apply_bounds_arr_dev(P(x), P(xo), N, N, N, None)),
    ("apply_bounds_rnlm (in place)", N, 16 * 2 * 9, lambda: lib.sfb_apply_bounds_rnlm_arr_dev(P(xr), P(xr), N, N, N, None)),
    ("reduce_M (complex)", Nm, 16 * n * n + 32 * r * r, lambda: lib.sfb_reduce_M_arr_dev(P(Mc), 1, Nm, Nm, *[P(t) for t in R4], None)),
    ("M_REG_arr", Nm, 8 * n * n + 72, lambda: lib.sfb_M_REG_arr_dev(P(eps), Nm, N, P(Mr), None)),
]
a4in = f64(81, N)
a6in = f64(729, Na6)
a2in = f64(9, N)
for t_ in (a2in, a4in, a6in):
    t_.normal_()
cases += [
    ("a2_to_nlm", N, 72 + 6 * 16, lambda: lib.sfb_ai_to_nlm_arr_dev(2, P(a2in), N, N, P(xo), N, None)),
    ("a4_to_nlm", N, 648 + 15 * 16, lambda: lib.sfb_ai_to_nlm_arr_dev(4, P(a4in), N, N, P(xo), N, None)),
    ("a6_to_nlm", Na6, 5832 + 28 * 16, lambda: lib.sfb_ai_to_nlm_arr_dev(6, P(a6in), Na6, Na6, P(xo), N, None)),
]
for name, Nn, bytes_node, fn in cases:
    for _ in range(3):
        _lib.check(fn())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    gbs = bytes_node * Nn / (ms * 1e-3) / 1e9
    print(json.dumps(dict(op=name, L=L, N=Nn, ms=round(ms, 4), per_s=round(Nn / (ms * 1e-3)), alg_bytes=bytes_node, gbs=round(gbs, 1), hbm_frac=round(gbs / HBM, 3))), flush=True)
