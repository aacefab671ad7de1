"""Tuning aid: time every compiled kernel variant of (L, terms, scheme) and check it against variant 0.
The tuning variants (build.py EXTRA) are only compiled with SFB_EXTRA_VARIANTS=1 python -m specfab_b200.build;
the default build holds variant 0, the RK4 default (100) and the full-form kernel (40)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import specfab_b200 as sf
from specfab_b200 import _lib


def run(L, N, terms, scheme, variants, steps=10, physical=True, reduced=False):
    lm, n = sf.init(L)
    g = torch.Generator(device="cuda").manual_seed(1)
    nlm = torch.zeros((n, N), dtype=torch.complex128, device="cuda")
    nlm[0] = 0.2820947917738781
    nlm[1:] = 1e-2 * torch.view_as_complex(torch.randn((n - 1, N, 2), dtype=torch.float64, device="cuda", generator=g))
    if physical:      # real-ODF symmetry n_l^-m = (-1)^m conj(n_l^m), Im n_l^0 = 0 (what the reduced kernels detect)
        j = 0
        for l in range(0, L + 1, 2):
            base = l * (l + 1) // 2
            nlm[base] = nlm[base].real.to(torch.complex128)
            for m in range(1, l + 1):
                nlm[base - m] = (-1) ** m * nlm[base + m].conj()
    stepf = sf.step_arr_dev
    if reduced:           # state in reduced form (rows m >= 0), step_rnlm_arr_dev
        rows = [l * (l + 1) // 2 + m for l in range(0, L + 1, 2) for m in range(0, l + 1)]
        nlm = nlm[torch.tensor(rows, device="cuda")].contiguous()
        stepf = sf.step_rnlm_arr_dev
    ug = torch.randn((3, 3, N), dtype=torch.float64, device="cuda", generator=g)
    tau = torch.randn((3, 3, N), dtype=torch.float64, device="cuda", generator=g)
    tau = (tau + tau.permute(1, 0, 2)) / 2
    kw = dict(dt=1e-3, Gamma0=4.0, Lambda=1.0, terms=terms, scheme=scheme)
    ref = None
    info = {(k["L"], k["ddrx"], k["variant"]): k for k in sf.build_info()["step_kernels"]}
    for v in variants:
        key = (L, int("ddrx" in terms), v)
        if key not in info:
            continue
        _lib.load().sfb_set_variant(v)
        out = torch.empty_like(nlm)
        try:
            for _ in range(3):
                stepf(nlm, ug, tau, out=out, **kw)
        except Exception as ex:
            print(json.dumps(dict(L=L, variant=v, error=str(ex)[:120])), flush=True)
            continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            stepf(nlm, ug, tau, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        k = info[key]
        nst = 4 if scheme == "rk4" else 1
        rate = N / (ms * 1e-3)
        print(json.dumps(dict(L=L, terms="+".join(terms), scheme=scheme, physical=physical, reduced=reduced, variant=v, roles=k["roles"], tile=k["tile"], ms=round(ms, 4),
                              rate=round(rate / 1e6, 1), tflops=round(2 * k["dfma_per_node_rhs"] * nst * rate / 1e12, 2), diff_vs_v0=err)), flush=True)
    _lib.load().sfb_set_variant(0)


if __name__ == "__main__":
    V = list(range(0, 40))
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if len(sys.argv) > 2:
        V = [int(x) for x in sys.argv[2].split(",")]
    if which == "one":      # one L N terms scheme variants [physical(1/0)] [reduced(0/1)]
        L, N, terms, scheme = int(sys.argv[2]), int(sys.argv[3]), tuple(sys.argv[4].split("+")), sys.argv[5]
        V = [int(x) for x in sys.argv[6].split(",")]
        run(L, N, terms, scheme, V, steps=10, physical=(len(sys.argv) <= 7 or sys.argv[7] == "1"), reduced=(len(sys.argv) > 8 and sys.argv[8] == "1"))
        sys.exit(0)
    if which == "w":        # windowed loop kernels (variants 60+) against the defaults, physical / general / reduced-form states
        for terms in (("lrot", "ddrx", "reg"), ("lrot", "reg")):
            for scheme in ("euler", "rk4"):
                run(8, 1_000_000, terms, scheme, V, steps=10)
        run(8, 200_000, ("lrot", "ddrx", "reg"), "euler", V, steps=5, physical=False)
        run(8, 200_000, ("lrot", "ddrx", "reg"), "rk4", V, steps=5, physical=False)
        run(8, 200_000, ("lrot", "reg"), "rk4", V, steps=5, physical=False)
        run(8, 1_000_003, ("lrot", "ddrx", "reg"), "euler", V, steps=5, reduced=True)
        run(8, 1_000_003, ("lrot", "ddrx", "reg"), "rk4", V, steps=5, reduced=True)
        run(8, 1_000_003, ("lrot", "reg"), "rk4", V, steps=5, reduced=True)
    if which == "8xr":      # headline kernels on reduced-form states
        run(8, 1_000_000, ("lrot", "reg"), "rk4", V, steps=20, reduced=True)
        run(8, 1_000_000, ("lrot", "reg"), "euler", V, steps=20, reduced=True)
    if which == "8x":       # headline kernels only (L=8, LROT+REG)
        run(8, 1_000_000, ("lrot", "reg"), "rk4", V, steps=20)
        run(8, 1_000_000, ("lrot", "reg"), "euler", V, steps=20)
    if which == "dd":       # DDRX kernels L = 6..12
        run(8, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V, steps=20)
        run(8, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V, steps=10)
        run(6, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V, steps=20)
        run(6, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V, steps=10)
        run(10, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V, steps=10)
        run(12, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V, steps=10)
    if which == "loop":     # reduced loop kernels
        run(12, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V, steps=10)
        run(20, 300_000, ("lrot", "ddrx", "cdrx", "reg"), "euler", V, steps=10)
    if which == "4r":       # two-lane reduced kernels (L = 6, 8 with DDRX)
        for L in (8, 6):
            run(L, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V, steps=20)
            run(L, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V, steps=10)
    if which == "exp":      # the experimental variants of build.py EXP (SFB_EXP_VARIANTS=1)
        for L, N in ((8, 1_000_000), (4, 2_000_000), (6, 1_000_000), (10, 1_000_000), (12, 500_000)):
            run(L, N, ("lrot", "reg"), "rk4", V, steps=20)
            run(L, N, ("lrot", "reg"), "euler", V, steps=20)
        for L in (4, 6):
            run(L, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V, steps=20)
            run(L, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V, steps=20)
    if which == "small":
        for L in (4, 6, 10):
            run(L, 1_000_000, ("lrot", "reg"), "rk4", V)
            run(L, 1_000_000, ("lrot", "reg"), "euler", V)
            run(L, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V)
            run(L, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V)
    if which == "ip":
        run(4, 1_000_000, ("lrot", "reg"), "rk4", V)
        run(4, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V)
        run(6, 1_000_000, ("lrot", "reg"), "rk4", V)
        run(10, 1_000_000, ("lrot", "reg"), "rk4", V)
        run(12, 500_000, ("lrot", "reg"), "rk4", V)
    if which == "r4":
        run(6, 1_000_000, ("lrot", "reg"), "rk4", V)
        run(6, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V)
        run(6, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V)
        run(8, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V)
        run(8, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V)
        run(10, 1_000_000, ("lrot", "reg"), "rk4", V)
        run(10, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V)
        run(12, 500_000, ("lrot", "reg"), "rk4", V)
    if which == "mid":
        run(8, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V)
        run(10, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V)
        run(10, 1_000_000, ("lrot", "reg"), "rk4", V)
        run(12, 500_000, ("lrot", "reg"), "rk4", V)
        run(12, 500_000, ("lrot", "reg"), "euler", V)
        for L in (14, 16, 18):
            run(L, 300_000, ("lrot", "ddrx", "cdrx", "reg"), "euler", V)
            run(L, 300_000, ("lrot", "reg"), "euler", V)
    if which == "8g":       # general complex states: the reduced kernels must take their in-kernel fallback
        run(8, 200_000, ("lrot", "reg"), "rk4", V, physical=False)
        run(8, 200_000, ("lrot", "ddrx", "reg"), "rk4", V, physical=False)
        run(8, 200_000, ("lrot", "ddrx", "reg"), "euler", V, physical=False)
    if which in ("all", "8"):
        run(8, 1_000_000, ("lrot", "reg"), "rk4", V)
        run(8, 1_000_000, ("lrot", "reg"), "euler", V)
        run(8, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V)
        run(8, 1_000_000, ("lrot", "ddrx", "reg"), "rk4", V)
    if which in ("all", "12"):
        run(12, 1_000_000, ("lrot", "ddrx", "reg"), "euler", V)
        run(12, 500_000, ("lrot", "reg"), "rk4", V)
    if which in ("all", "20"):
        run(20, 300_000, ("lrot", "ddrx", "cdrx", "reg"), "euler", V)
        run(20, 300_000, ("lrot", "reg"), "euler", V)
    if which in ("all", "4"):
        run(4, 2_000_000, ("lrot", "reg"), "rk4", V)
        run(4, 2_000_000, ("lrot", "ddrx", "reg"), "euler", V)
