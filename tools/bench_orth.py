"""Time Eij_orthotropic_arr_dev on resident states (node-updates/s), both third-axis branches."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import specfab_b200 as sf
from util import random_states

L = 8
sf.init(L)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
q = [torch.from_numpy(np.ascontiguousarray(random_states(L, N, s, True, 0.35).T)).cuda() for s in (1, 2, 3)]
rng = np.random.default_rng(0)
Q = np.linalg.qr(rng.standard_normal((N, 3, 3)))[0]
e = [torch.from_numpy(np.ascontiguousarray(Q[:, :, i].T)).cuda() for i in range(3)]
out = torch.empty((6, N), dtype=torch.float64, device="cuda")
G = (1.0, 1.0, 1.0, 1.0, 1.0, 10.0)
for name, q3 in (("derived", None), ("given", q[2])):
    for _ in range(3):
        sf.Eij_orthotropic_arr_dev(q[0], q[1], q3, e[0], e[1], e[2], G, 0.0, 1, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        sf.Eij_orthotropic_arr_dev(q[0], q[1], q3, e[0], e[1], e[2], G, 0.0, 1, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("Eij_orthotropic %-8s N=%d  %.3f ms  %.3e nodes/s" % (name, N, ms, N / ms * 1e3), flush=True)
