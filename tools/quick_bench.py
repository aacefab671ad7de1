"""Quick device-resident timing of the step kernels (development aid; bench.py is the contract)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specfab_b200 as sf


def bench(L, N, terms, scheme, steps=20, warm=3):
    lm, n = sf.init(L)
    g = torch.Generator(device="cuda").manual_seed(1)
    nlm = torch.zeros((n, N), dtype=torch.complex128, device="cuda")
    nlm[0] = 0.2820947917738781
    nlm[1:] = 1e-3 * torch.view_as_complex(torch.randn((n - 1, N, 2), dtype=torch.float64, device="cuda", generator=g))
    ug = torch.randn((3, 3, N), dtype=torch.float64, device="cuda", generator=g)
    tau = torch.randn((3, 3, N), dtype=torch.float64, device="cuda", generator=g)
    tau = (tau + tau.permute(1, 0, 2)) / 2
    out = torch.empty_like(nlm)
    kw = dict(dt=1e-3, Gamma0=4.0, Lambda=1.0, terms=terms, scheme=scheme)
    for _ in range(warm):
        sf.step_arr_dev(nlm, ug, tau, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sf.step_arr_dev(nlm, ug, tau, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    info = [k for k in sf.build_info()["step_kernels"] if k["L"] == L and k["ddrx"] == int("ddrx" in terms)][0]
    nst = 4 if scheme == "rk4" else 1
    rate = N / (ms * 1e-3)
    return dict(L=L, N=N, terms="+".join(terms), scheme=scheme, ms=round(ms, 4), node_updates_per_s=rate,
                gbs_alg=round((32 * n + 72 + (72 if "ddrx" in terms else 0)) * rate / 1e9, 1),
                dfma_tflops=round(2 * info["dfma_per_node_rhs"] * nst * rate / 1e12, 2))


if __name__ == "__main__":
    cases = [(8, 1_000_000, ("lrot", "reg"), "euler"), (8, 1_000_000, ("lrot", "reg"), "rk4"),
             (8, 1_000_000, ("lrot", "ddrx", "reg"), "euler"), (12, 1_000_000, ("lrot", "ddrx", "reg"), "euler"),
             (20, 200_000, ("lrot", "ddrx", "cdrx", "reg"), "euler"), (4, 1_000_000, ("lrot", "reg"), "euler"),
             (12, 1_000_000, ("lrot", "reg"), "rk4"), (20, 200_000, ("lrot", "reg"), "rk4")]
    for c in cases:
        print(json.dumps(bench(*c)), flush=True)
