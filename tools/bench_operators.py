"""Time the batched reduced-operator export (what an implicit FE fabric solve consumes every step) on
resident inputs: M_LROT_reduced_arr_dev / M_DDRX_reduced_arr_dev -> Mrr, Mri, Mir, Mii (N, r, r) real(8) each.
Prints nodes/s and the output bandwidth (4 r^2 8 B per node are written) against the measured HBM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch
import specfab_b200 as sf
from specfab_b200 import _lib
from util import random_states, random_tau, random_ugrad

try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
lib = _lib.load()
for L, N in ((8, 100_000), (12, 40_000), (4, 400_000)):
    lm, n = sf.init(L)
    r = sf.rnlm_len()
    x = torch.from_numpy(np.ascontiguousarray(random_states(L, N, 1, True, 0.35).T)).cuda()
    u = random_ugrad(N, 2)
    D, W = (u + u.transpose(0, 2, 1)) / 2, (u - u.transpose(0, 2, 1)) / 2
    lay = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda()       # (3,3,N)
    Dd, Wd, Td = lay(D), lay(W), lay(random_tau(N, 3))
    outs = [torch.empty((r, r, N), dtype=torch.float64, device="cuda") for _ in range(4)]
    ptrs = [o.data_ptr() for o in outs]

    def lrot():
        _lib.check(lib.sfb_M_LROT_reduced_arr_dev(Dd.data_ptr(), Wd.data_ptr(), N, N, 1.0, 0.0, *ptrs, None))

    def ddrx():
        _lib.check(lib.sfb_M_DDRX_reduced_arr_dev(x.data_ptr(), N, Td.data_ptr(), N, N, 0, *ptrs, None))

    Nd = max(N // 4, 1000)                                      # dense (N, n, n) complex(8) output: 16 n^2 B per node
    Md = torch.empty((n, n, Nd), dtype=torch.complex128, device="cuda")

    def lrot_dense():
        _lib.check(lib.sfb_M_LROT_arr_dev(Dd.data_ptr(), Wd.data_ptr(), Nd, N, 1.0, 0.0, Md.data_ptr(), None))

    def ddrx_dense():
        _lib.check(lib.sfb_M_DDRX_arr_dev(x.data_ptr(), N, Td.data_ptr(), Nd, N, Md.data_ptr(), None))

    for name, fn in (("M_LROT_reduced", lrot), ("M_DDRX_reduced", ddrx), ("M_LROT_dense", lrot_dense), ("M_DDRX_dense", ddrx_dense)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        dense = name.endswith("dense")
        Nn = Nd if dense else N
        gbs = (16 * n * n if dense else 4 * r * r * 8) * Nn / (ms * 1e-3) / 1e9
        print(json.dumps(dict(op=name, L=L, N=Nn, ms=round(ms, 4), nodes_per_s=round(Nn / (ms * 1e-3)), out_gbs=round(gbs, 1), hbm_frac=round(gbs / HBM, 3))), flush=True)
