import sys, os
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import numpy as np, torch
import specfab_b200 as sf
from util import random_states
sf.init(8)
N = 2000000
x = torch.from_numpy(np.ascontiguousarray(random_states(8, N, 1, True, 0.35).T)).cuda()
out = torch.empty((6, N), dtype=torch.float64, device="cuda")
for _ in range(3):
    sf.Eij_eigenframe_arr_dev(x, (1.0, 1e3), 0.0125, 1, out=out)
torch.cuda.synchronize()
