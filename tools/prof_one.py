"""Run a few launches of one step-kernel configuration (for ncu).  usage: prof_one.py L N terms scheme [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specfab_b200 as sf

L, N, terms, scheme = int(sys.argv[1]), int(sys.argv[2]), tuple(sys.argv[3].split("+")), sys.argv[4]
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
variant = int(sys.argv[6]) if len(sys.argv) > 6 else 0
lm, n = sf.init(L)
from specfab_b200 import _lib
_lib.load().sfb_set_variant(variant)
g = torch.Generator(device="cuda").manual_seed(1)
nlm = torch.zeros((n, N), dtype=torch.complex128, device="cuda")
nlm[0] = 0.2820947917738781
nlm[1:] = 1e-3 * torch.view_as_complex(torch.randn((n - 1, N, 2), dtype=torch.float64, device="cuda", generator=g))
if os.environ.get("SFB_GENERAL_STATES", "0") != "1":      # real-ODF symmetry (what the reduced kernels detect)
    for l in range(0, L + 1, 2):
        base = l * (l + 1) // 2
        nlm[base] = nlm[base].real.to(torch.complex128)
        for m in range(1, l + 1):
            nlm[base - m] = (-1) ** m * nlm[base + m].conj()
ug = torch.randn((3, 3, N), dtype=torch.float64, device="cuda", generator=g)
tau = torch.randn((3, 3, N), dtype=torch.float64, device="cuda", generator=g)
tau = (tau + tau.permute(1, 0, 2)) / 2
stepf = sf.step_arr_dev
if os.environ.get("SFB_RNLM", "0") == "1":               # state in reduced form (rows m >= 0): step_rnlm_arr_dev
    rows = [l * (l + 1) // 2 + m for l in range(0, L + 1, 2) for m in range(0, l + 1)]
    nlm = nlm[torch.tensor(rows, device="cuda")].contiguous()
    stepf = sf.step_rnlm_arr_dev
out = torch.empty_like(nlm)
for _ in range(steps):
    stepf(nlm, ug, tau, out=out, dt=1e-3, Gamma0=4.0, Lambda=1.0, terms=terms, scheme=scheme)
torch.cuda.synchronize()
