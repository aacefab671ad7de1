#!/usr/bin/env python3
"""bench.py -- fabric node-updates/s of the B200 fabric-evolution engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config 2|3|4|5] [--impl reference]

One "step" = one time step of every node of the synthetic field (SURVEY.md 8d inputs).  Default
workload = BASELINE config 2 (configs[1]): Eulerian field of 1e6 nodes per GPU, L=8, LROT+REG,
RK4, synthetic per-node velocity gradients.  N>1: one process per GPU (torchrun), contiguous node
ranges, no collective while stepping (weak scaling: 1e6 nodes per GPU); the only NCCL traffic is
the barrier and the max-over-ranks reduction of the device time.

Prints ONE JSON line (rank 0).  `value` = node-updates/s with the state resident in HBM, timed
with CUDA events on the launch stream; `e2e` = the same metric through the host-pointer C-ABI
call (pinned host buffers, H2D + D2H inside the timed region); `roofline` for the step kernel
(FP64 CUDA-core bound: peak = measured DFMA-chain peak, profiles/r01_fp64_peak.json; the HBM
fraction against MEASURED_PEAKS.json is reported next to it); `cpu_baseline` = the dense C
restatement of the reference algorithm (oracle/, gfortran is unavailable) on the host cores.
--impl reference runs that CPU restatement instead of the GPU engine.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np

DT = -np.log(0.02) / 1000          # strain -0.98 in 1000 steps (demo/fabric-evolution/...latrot.py:34)
CONFIGS = {
    # id: (L, nodes per GPU, terms, scheme, with Eij outputs, description)
    2: (8, 1_000_000, ("lrot", "reg"), "rk4", False, "cfg2: 1e6 nodes/GPU, L=8, LROT+REG, RK4, synthetic ugrad"),
    3: (12, 1_250_000, ("lrot", "ddrx", "reg"), "euler", False, "cfg3: 1e7 nodes over 8 GPUs (1.25e6/GPU), L=12, LROT+DDRX+REG, Euler"),
    4: (20, 1_000_000, ("lrot", "ddrx", "cdrx", "reg"), "euler", False, "cfg4: 1e6 nodes/GPU, L=20, LROT+CDRX+DDRX+REG, Euler"),
    5: (8, 1_250_000, ("lrot", "ddrx", "reg"), "euler", True, "cfg5: 1e7 nodes over 8 GPUs (1.25e6/GPU), L=8, step + a2/eig/Eij per node"),
}
GRAIN, ALPHA = (1.0, 1e3), 0.0125   # ice 'linear' (src/specfabpy/constants.py:10)


def synth_forcing(N, seed):
    """ugrad = traceless standard normal scaled to ||D||_F = sqrt(1.5); tau = traceless
    symmetric normal scaled to ||tau||_F = 1.  Fortran (N,3,3) order = numpy (3,3,N)[k,i,p]."""
    rng = np.random.default_rng(seed)
    u = rng.standard_normal((N, 3, 3))
    u -= np.eye(3)[None] * (np.trace(u, axis1=1, axis2=2) / 3)[:, None, None]
    D = (u + u.transpose(0, 2, 1)) / 2
    u *= (np.sqrt(1.5) / np.sqrt((D ** 2).sum(axis=(1, 2))))[:, None, None]
    a = rng.standard_normal((N, 3, 3))
    t = (a + a.transpose(0, 2, 1)) / 2
    t -= np.eye(3)[None] * (np.trace(t, axis1=1, axis2=2) / 3)[:, None, None]
    t /= np.sqrt((t ** 2).sum(axis=(1, 2)))[:, None, None]
    return u, t


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa(torch, local):
    """Pin this rank's host threads (and, by first touch, its pinned staging buffers) to the NUMA node of its GPU -- the
    host-pointer leg moves 1.5 KB per node-step across PCIe and every rank does so at once.  Best effort: any failure (no
    sysfs entry, cpuset without local cores) leaves the affinity alone.  Returns a short description for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d has no allowed cpus" % node
        os.sched_setaffinity(0, cpus)
        return "bound to numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as ex:   # noqa
        return "not bound (%s)" % type(ex).__name__


def fp64_peak():
    p = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    try:
        return json.load(open(p))["fp64_tflops_sustained"], "measured DFMA-chain peak (profiles/r01_fp64_peak.json, tools/fp64_peak.cu)"
    except Exception:
        return 37.0, "nominal FP64 (fallback; profiles/r01_fp64_peak.json missing)"


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def alg_bytes_per_node_step(n, terms, eij):
    b = 32 * n + 72 + (72 if "ddrx" in terms else 0)
    if eij:
        b += 240 + 72 + 24 + 48                                # Eij kernel: 15 coef rows in, ei + lami + Eij out
    return b


# ----------------------------------------------------------------------------------------------
# CPU arm: dense C restatement of the reference algorithm (oracle/specfab_oracle.c)
# ----------------------------------------------------------------------------------------------
def cpu_rate(cfg, nodes, reps=1):
    import oracle_c as oc
    L, _, terms, scheme, eij, _ = CONFIGS[cfg]
    n = oc.init(L)
    u, t = synth_forcing(nodes, 20260817)
    x = np.zeros((nodes, n), dtype=np.complex128)
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    kw = dict(dt=DT, Gamma0=4.0, Lambda=1.0, use_lrot="lrot" in terms, use_ddrx="ddrx" in terms, use_cdrx="cdrx" in terms,
              use_reg="reg" in terms, scheme=scheme)
    x = oc.step_batch(x, u, t, nsteps=2, **kw)     # leave isotropy so that all coefficients are non-zero
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        x = oc.step_batch(x, u, t, nsteps=1, **kw)
        best = max(best, nodes / (time.perf_counter() - t0))
    return best, oc.num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores (C port: no Fortran compiler)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    L, npg, terms, scheme, eij, desc = CONFIGS[cfg]
    probe, cores = cpu_rate(cfg, 2000)
    sample = int(max(2000, min(npg, probe * 2.0)))           # ~2 s of CPU work per step
    import oracle_c as oc
    n = oc.init(L)
    u, t = synth_forcing(sample, 20260817)
    x = np.zeros((sample, n), dtype=np.complex128)
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    kw = dict(dt=DT, Gamma0=4.0, Lambda=1.0, use_lrot="lrot" in terms, use_ddrx="ddrx" in terms, use_cdrx="cdrx" in terms,
              use_reg="reg" in terms, scheme=scheme)
    for _ in range(args.warmup):
        x = oc.step_batch(x, u, t, nsteps=1, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        x = oc.step_batch(x, u, t, nsteps=1, **kw)
    el = time.perf_counter() - t0
    val = sample * args.steps / el
    line = {"impl": "reference", "metric": "fabric node-updates/s", "value": val, "unit": "node-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "L": L, "terms": "+".join(terms), "scheme": scheme},
            "cpu_baseline": {"value": val, "unit": "node-updates/s", "cores": cores, "kind": "port",
                             "sample": "%d nodes per step (bounded sample of the workload); dense per-node operator build + matvec, "
                                       "C restatement of src/dynamics.f90:94-96,108 (no Fortran compiler in the image)" % sample},
            "e2e": {"value": val, "unit": "node-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def rnlm_rows(L):
    """rows of nlm that make up the reduced form (m >= 0; src/reducedform.f90:160-187)"""
    lm = [(l, m) for l in range(0, L + 1, 2) for m in range(-l, l + 1)]
    return [lm.index((l, m)) for l in range(0, L + 1, 2) for m in range(0, l + 1)]


def time_config(sf, torch, cfg, N, rank, steps, warmup, sampler=None, reduced=False, scheme=None):
    """device-resident timing of one config; returns (ms_per_step_local, launches_per_step, n).
    reduced=True: the state is kept in reduced form (rnlm) and stepped with step_rnlm_arr_dev."""
    L, _, terms, scheme0, eij, _ = CONFIGS[cfg]
    scheme = scheme or scheme0
    lm, n = sf.init(L)
    u, t = synth_forcing(N, 20260817 + rank)
    ug = torch.from_numpy(np.ascontiguousarray(u.transpose(2, 1, 0))).cuda()     # (3,3,N): [k,i,p]
    tau = torch.from_numpy(np.ascontiguousarray(t.transpose(2, 1, 0))).cuda() if "ddrx" in terms else None
    nlm = torch.zeros((n, N), dtype=torch.complex128, device="cuda")
    nlm[0] = 1 / np.sqrt(4 * np.pi)
    kw = dict(dt=DT, Gamma0=4.0, Lambda=1.0, terms=terms, scheme=scheme)
    # 50 spin-up steps from isotropy so that every coefficient is non-zero
    for _ in range(50):
        sf.step_arr_dev(nlm, ug, tau, dt=DT, Gamma0=4.0, Lambda=1.0, terms=terms, scheme="euler")
    stepf = sf.step_arr_dev
    if reduced:
        nlm = nlm[torch.tensor(rnlm_rows(L), device="cuda")].contiguous()
        stepf = sf.step_rnlm_arr_dev
        n = nlm.shape[0]
    eout = torch.empty((6, N), dtype=torch.float64, device="cuda") if eij else None
    eiv = torch.empty((3, 3, N), dtype=torch.float64, device="cuda") if eij else None
    lam = torch.empty((3, N), dtype=torch.float64, device="cuda") if eij else None

    def one():
        stepf(nlm, ug, tau, **kw)
        if eij:
            sf.Eij_eigenframe_arr_dev(nlm, GRAIN, ALPHA, 1, out=eout, ei=eiv, lami=lam)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    finite = bool(torch.isfinite(torch.view_as_real(nlm)).all().item())
    del nlm, ug, tau, eout, eiv, lam
    torch.cuda.empty_cache()
    return ms, (2 if eij else 1), n, finite


def time_e2e(sf, torch, cfg, N, rank, steps, warmup, reduced=False):
    """the same step through the host-pointer C-ABI call: pinned host buffers, H2D of state + forcing and
    D2H of the new state inside the timed region, every step.  reduced=True: sfb_step_rnlm_arr on reduced-form states."""
    L, _, terms, scheme, eij, _ = CONFIGS[cfg]
    lm, n = sf.init(L)
    hstep = sf.step_arr
    if reduced:
        n = sf.rnlm_len()
        hstep = sf.step_rnlm_arr
    u, t = synth_forcing(N, 20260817 + rank)

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory().numpy()

    x = pinned((n, N), torch.complex128).T        # Fortran-ordered (N, n) view of pinned memory
    x[:] = 0
    x[:, 0] = 1 / np.sqrt(4 * np.pi)
    ugp = pinned((3, 3, N), torch.float64).transpose(2, 1, 0)
    ugp[:] = u
    tp = None
    if "ddrx" in terms:
        tp = pinned((3, 3, N), torch.float64).transpose(2, 1, 0)
        tp[:] = t
    kw = dict(dt=DT, Gamma0=4.0, Lambda=1.0, terms=terms, scheme=scheme)
    x = hstep(x, ugp, tp, **dict(kw, scheme="euler", nsteps=20))
    xin = pinned((n, N), torch.complex128).T
    xin[:] = x
    xout = pinned((n, N), torch.complex128).T
    for _ in range(warmup):
        hstep(xin, ugp, tp, out=xout, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = hstep(xin, ugp, tp, out=xout, **kw)
        if eij:
            sf.Eij_eigenframe_arr(out, GRAIN, ALPHA, 1)
    el = time.perf_counter() - t0
    h2d = N * (16 * n + 72 + (72 if tp is not None else 0)) + (N * 240 if eij else 0)
    d2h = N * 16 * n + (N * (48 + 72 + 24 + 4) if eij else 0)
    return el / steps, h2d, d2h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--nodes", type=int, default=None, help="nodes per GPU (default: the config's)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs / cpu baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import specfab_b200 as sf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(torch, local) if world > 1 else "single rank: not bound"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tns = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    cfg = args.config
    L, npg, terms, scheme, eij, desc = CONFIGS[cfg]
    N = args.nodes or npg
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ms, lps, n, finite = time_config(sf, torch, cfg, N, rank, args.steps, args.warmup, sampler)
    clocks = sampler.result() if sampler else None
    barrier()
    ms = max_over_ranks(ms)
    value = N * world / (ms * 1e-3)

    # end-to-end through the host-pointer C ABI (fewer steps: PCIe bound)
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    sec, h2d, d2h = time_e2e(sf, torch, cfg, N, rank, e2e_steps, 1)
    barrier()
    sec = max_over_ranks(sec)
    e2e_val = N * world / sec

    info = [k for k in sf.build_info()["step_kernels"] if k["L"] == L and k["ddrx"] == int("ddrx" in terms) and k["variant"] == 0][0]
    nst = 4 if scheme == "rk4" else 1
    per_gpu_rate = N / (ms * 1e-3)
    fpk, fsrc = fp64_peak()
    hpk, hsrc = hbm_peak()
    # FP64 work per node-step, in FP64-pipe instruction slots (a DFMA, DMUL or DADD occupies the pipe alike; the measured
    # peak is a DFMA chain = 2 flop per slot).  Counted conservatively: min(instructions the kernel EXECUTES -- ncu,
    # profiles/fp64_ops.json; the compiler drops e.g. the imaginary parts of the m = 0 rows --, the code generator's count
    # of the factorised algorithm -- which excludes the zero padding the table-driven loop kernels execute).
    nominal = info["dfma_per_node_rhs"] * nst
    ops = None
    try:
        ops = json.load(open(os.path.join(ROOT, "profiles", "fp64_ops.json"))).get({2: "cfg2", 3: "cfg3", 4: "cfg4", 5: "cfg5_step"}[cfg])
    except Exception:
        pass
    executed = (ops["dfma"] + ops["dmul"] + ops["dadd"]) if ops else None
    slots = min(executed, nominal) if executed else nominal
    ach_tf = 2.0 * slots * per_gpu_rate / 1e12
    ach_gb = alg_bytes_per_node_step(n, terms, eij) * per_gpu_rate / 1e9
    fp_frac, hbm_frac = ach_tf / fpk, ach_gb / hpk
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr.get("cfg%d" % cfg)
    except Exception:
        pass
    if fp_frac >= hbm_frac:
        roof = {"bound": "fp64", "achieved": ach_tf, "peak": fpk, "unit": "TFLOP/s", "frac": fp_frac, "peak_source": fsrc}
    else:
        roof = {"bound": "hbm", "achieved": ach_gb, "peak": hpk, "unit": "GB/s", "frac": hbm_frac, "peak_source": hsrc}
    reduced = info["roles"] >= 1 and info["tile"] == 32 and any(
        k["L"] == L and k["ddrx"] == info["ddrx"] and k["variant"] == 40 for k in sf.build_info()["step_kernels"])
    roof.update({"traffic": traffic,
                 "kernel": "%s (L=%d, %s, %s)" % ("step_kernel_r" if reduced else "step_kernel", L, "+".join(terms), scheme),
                 "form": ("reduced: only the rows m >= 0 are computed (real-ODF symmetry of every 32-node tile verified in the kernel, "
                          "full-form fallback otherwise; SURVEY.md 8d: flop counts scaled accordingly)") if reduced else "full",
                 "fp64_accounting": "achieved = 2 flop x FP64-pipe instruction slots per node-step x node rate (DFMA-equivalent: DFMA, DMUL and "
                                    "DADD occupy the pipe alike and the measured peak is a DFMA chain); slots = min(executed per ncu, "
                                    "code generator's count); cross-check: ncu sm__pipe_fp64_cycles_active of the same kernel = "
                                    "%s %% (profiles/fp64_ops.json)" % (ops["pipe_fp64_pct"] if ops else "n/a"),
                 "fp64_slots_per_node_step": slots,
                 "fp64_ops_executed_per_node_step": ({"dfma": ops["dfma"], "dmul": ops["dmul"], "dadd": ops["dadd"]} if ops else None),
                 "flops_executed_per_node_step": (2 * ops["dfma"] + ops["dmul"] + ops["dadd"]) if ops else None,
                 "fp64_slots_codegen_per_node_step": nominal,
                 "alg_bytes_per_node_step": alg_bytes_per_node_step(n, terms, eij),
                 "hbm": {"achieved": ach_gb, "peak": hpk, "unit": "GB/s", "frac": hbm_frac, "peak_source": hsrc},
                 "fp64": {"achieved": ach_tf, "peak": fpk, "unit": "TFLOP/s", "frac": fp_frac, "peak_source": fsrc},
                 "note": "FP64 CUDA-core (DFMA) bound path: tensor cores do not apply (DESIGN.md); the governing bound is reported"})

    extra = {}
    cpu = None
    if rank == 0 and world == 1 and not args.no_extra:
        for c in sorted(CONFIGS):
            if c == cfg:
                continue
            try:
                Lc, npc, tc, sc, ec, dc = CONFIGS[c]
                Nc = min(npc, 1_000_000 if Lc >= 12 else npc)
                m2, l2, n2, fin2 = time_config(sf, torch, c, Nc, rank, 10, 3)
                extra["cfg%d" % c] = {"workload": dc, "nodes": Nc, "ms_per_step": m2, "node_updates_per_s": Nc / (m2 * 1e-3),
                                      "finite": fin2}
            except Exception as ex:   # noqa
                extra["cfg%d" % c] = {"error": str(ex)[:200]}
        # the headline workload on reduced-form states (rows m >= 0 only; the FE couplers' state representation)
        try:
            r = (L + 2) ** 2 // 4
            rr = {}
            for sc in ("rk4", "euler"):
                m3, _, _, fin3 = time_config(sf, torch, cfg, N, rank, 10, 3, reduced=True, scheme=sc)
                rr[sc] = {"ms_per_step": m3, "node_updates_per_s": N / (m3 * 1e-3), "alg_bytes_per_node_step": 32 * r + 72,
                          "hbm_gbs_alg": (32 * r + 72) * N / (m3 * 1e-3) / 1e9, "hbm_frac": (32 * r + 72) * N / (m3 * 1e-3) / 1e9 / hpk,
                          "fp64_frac": 2.0 * (slots / nst) * (4 if sc == "rk4" else 1) * N / (m3 * 1e-3) / 1e12 / fpk, "finite": fin3}
            if not eij:
                sec3, h3, d3 = time_e2e(sf, torch, cfg, N, rank, 3, 1, reduced=True)
                rr["e2e"] = {"value": N / sec3, "unit": "node-updates/s", "h2d_bytes_per_step": h3, "d2h_bytes_per_step": d3,
                             "api": "sfb_step_rnlm_arr (host pointers)"}
            rr["workload"] = "%s, state in reduced form (rnlm: %d of %d coefficient rows), sfb_step_rnlm_arr(_dev)" % (desc, r, n)
            extra["rnlm"] = rr
        except Exception as ex:   # noqa
            extra["rnlm"] = {"error": str(ex)[:200]}
        # BASELINE config 5 (FE coupling: step + eigenframe + Eij per node) on a field kept in reduced form
        try:
            L5, n5, t5, s5, _, d5 = CONFIGS[5]
            sf.init(L5)
            u5, tt5 = synth_forcing(n5, 20260817 + rank)
            ug5 = torch.from_numpy(np.ascontiguousarray(u5.transpose(2, 1, 0))).cuda()
            ta5 = torch.from_numpy(np.ascontiguousarray(tt5.transpose(2, 1, 0))).cuda()
            st5 = torch.zeros((sf.rnlm_len(), n5), dtype=torch.complex128, device="cuda")
            st5[0] = 1 / np.sqrt(4 * np.pi)
            kw5 = dict(dt=DT, Gamma0=4.0, Lambda=1.0, terms=t5, scheme=s5)
            for _ in range(50):
                sf.step_rnlm_arr_dev(st5, ug5, ta5, **kw5)
            for _ in range(3):
                sf.step_moments_Eij_rnlm_arr_dev(st5, ug5, ta5, GRAIN, ALPHA, 1, want_frame=True, **kw5)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                sf.step_moments_Eij_rnlm_arr_dev(st5, ug5, ta5, GRAIN, ALPHA, 1, want_frame=True, **kw5)
            b.record()
            torch.cuda.synchronize()
            m5 = a.elapsed_time(b) / 10
            extra["cfg5_rnlm"] = {"workload": d5 + ", state in reduced form (sfb_step_moments_Eij_rnlm_arr_dev)", "nodes": n5, "ms_per_step": m5,
                                  "node_updates_per_s": n5 / (m5 * 1e-3), "finite": bool(torch.isfinite(torch.view_as_real(st5)).all().item())}
            del st5, ug5, ta5
            sf.init(L)
        except Exception as ex:   # noqa
            extra["cfg5_rnlm"] = {"error": str(ex)[:200]}
        # stand-alone Eij evals/s (a2 -> eigenframe -> 6 Sachs/Taylor factors per node)
        try:
            lm, n8 = sf.init(8)
            Ne = 4_000_000
            st = torch.zeros((n8, Ne), dtype=torch.complex128, device="cuda")
            st[0] = 1 / np.sqrt(4 * np.pi)
            ugq, _ = synth_forcing(Ne, 7)
            ugd = torch.from_numpy(np.ascontiguousarray(ugq.transpose(2, 1, 0))).cuda()
            for _ in range(60):
                sf.step_arr_dev(st, ugd, None, dt=DT, terms=("lrot", "reg"))
            eo = torch.empty((6, Ne), dtype=torch.float64, device="cuda")
            for _ in range(3):
                sf.Eij_eigenframe_arr_dev(st, GRAIN, ALPHA, 1, out=eo)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                sf.Eij_eigenframe_arr_dev(st, GRAIN, ALPHA, 1, out=eo)
            b.record()
            torch.cuda.synchronize()
            me = a.elapsed_time(b) / 10
            extra["eij"] = {"workload": "stand-alone a2 -> eigenframe -> Eij_tranisotropic, 4e6 nodes, L=8", "ms": me,
                            "eij_evals_per_s": Ne / (me * 1e-3), "hbm_gbs_alg": 288 * Ne / (me * 1e-3) / 1e9}
            try:
                eo_ = json.load(open(os.path.join(ROOT, "profiles", "fp64_ops.json")))["eij"]
                sl = eo_["dfma"] + eo_["dmul"] + eo_["dadd"]
                extra["eij"].update({"fp64_slots_per_eval": sl, "fp64_frac": 2.0 * sl * Ne / (me * 1e-3) / 1e12 / fpk})
            except Exception:
                pass
            del st, ugd, eo
        except Exception as ex:   # noqa
            extra["eij"] = {"error": str(ex)[:200]}
        try:
            probe, cores = cpu_rate(cfg, 2000)
            sample = int(max(2000, min(N, probe * 12.0)))          # ~12 s of CPU work
            rate, cores = cpu_rate(cfg, sample)
            cpu = {"value": rate, "unit": "node-updates/s", "cores": cores, "kind": "port",
                   "sample": "%d nodes x 1 step of the same workload; dense per-node operator build + matvec "
                             "(C restatement of the reference algorithm, OpenMP over nodes; gfortran unavailable)" % sample}
        except Exception as ex:   # noqa
            cpu = {"value": None, "unit": "node-updates/s", "cores": 0, "kind": "port", "sample": "failed: %s" % str(ex)[:160]}

    if rank == 0:
        line = {"metric": "fabric node-updates/s", "value": value, "unit": "node-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "L": L, "nlm_len": n, "nodes_per_gpu": N, "terms": "+".join(terms), "scheme": scheme,
                           "dt": DT, "l2": "inputs exceed L2 (state %.0f MB per GPU vs 126 MB L2); no flush needed" % (N * n * 16 / 1e6),
                           "sharding": "contiguous node ranges, no collective while stepping"},
                "e2e": {"value": e2e_val, "unit": "node-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "api": "sfb_step_arr (host pointers, pinned buffers, chunked H2D|kernel|D2H pipeline)", "host_affinity": numa},
                "gpu_launches": lps * args.steps, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "finite": finite, "other_configs": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
