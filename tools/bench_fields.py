"""Time the per-node field evaluations on resident states (evaluations/s)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import specfab_b200 as sf
from specfab_b200 import _lib
from util import random_states, random_tau

L = 8
lm, n = sf.init(L)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
x = torch.from_numpy(np.ascontiguousarray(random_states(L, N, 1, True, 0.35).T)).cuda()
eps = torch.from_numpy(np.ascontiguousarray(random_tau(N, 2).reshape(N, 9, order="F").T)).cuda()     # (9, N): plane i+3j
out6 = torch.empty((6, N), dtype=torch.float64, device="cuda")
out1 = torch.empty(N, dtype=torch.float64, device="cuda")
G = np.array([1.0, 1e3])
lib = _lib.load()


def timeit(name, fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-34s N=%d  %8.3f ms  %.3e evals/s" % (name, N, ms, N / ms * 1e3), flush=True)


for ng in (1, 3):
    timeit("Eij_eigenframe n'=%d" % ng, lambda: sf.Eij_eigenframe_arr_dev(x, G, 0.0125, ng, out=out6))
    timeit("E_CAFFE n'=%d" % ng, lambda: _lib.check(lib.sfb_E_CAFFE_arr_dev(x.data_ptr(), N, N, eps.data_ptr(), 0.1, 10.0, ng, out1.data_ptr(), None)))
timeit("pfJ", lambda: _lib.check(lib.sfb_pfJ_arr_dev(x.data_ptr(), N, N, L, out1.data_ptr(), None)))
