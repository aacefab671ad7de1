"""Build libspecfab_b200.so (sm_100a) in-tree:  python -m specfab_b200.build [-j N] [--L 8,12]

1. generate the straight-line operator-apply code for every (L, term set)  (codegen/emit_step.py)
2. nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo each translation unit (parallel)
3. link specfab_b200/libspecfab_b200.so  (C ABI declared in include/specfab_b200.h)

nvcc cross-compiles without a GPU.  Generated sources go to specfab_b200/csrc/gen (git-ignored).
"""
import argparse
import concurrent.futures as cf
import hashlib
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "gen")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libspecfab_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-I", CSRC,
          "-I", os.path.join(HERE, "..", "include")]

ALL_L = [4, 6, 8, 10, 12, 14, 16, 18, 20]

# tuning table: (L, ddrx) -> (roles R, tile nodes TN, min CTAs/SM for __launch_bounds__)
# variant 0 is the default; extra variants (EXTRA) are selectable with sfb_set_variant() for tuning runs
EXTRA = {
    # (L, ddrx): [(variant id, R, TN, MINB, const_mode, sync)]
    #   R >= 1: two-lane straight-line kernel with R warp roles; R = 0: four-lane straight-line kernel;
    #   R = -1: persistent four-lane kernel (streaming TMA refill); R = -2: table-driven loop kernel (+rN roles, +chN rows/chunk)
    (8, 0): [(10, 0, 32, 3, "imm+w", True), (20, -1, 96, 1, "imm+w", True), (30, -2, 32, 3, "imm+ch2+r2", False),
             (11, 2, 16, 7, "imm", False), (12, 2, 16, 6, "imm", False), (13, 2, 16, 5, "imm", False), (14, 2, 32, 3, "imm", False),
             (15, 3, 16, 7, "imm", False), (16, 2, 16, 7, "imm+c6", False),
             (25, -3, 32, 7, "imm+r2", False), (26, -3, 32, 7, "imm+r3", False), (27, -3, 32, 7, "imm+c4", False), (28, -3, 32, 7, "imm+c8", False),
             (31, -3, 32, 9, "imm+ip", False), (32, -3, 32, 10, "imm+ip", False), (33, -3, 32, 11, "imm+ip", False), (34, -3, 32, 12, "imm+ip", False),
             (35, -3, 32, 8, "imm+ip", False)],
    (8, 1): [(1, 1, 32, 3, "imm", False), (20, -1, 64, 1, "imm+w", True), (30, -2, 32, 3, "imm+ch2+r2", False),
             (11, 0, 64, 2, "imm", False), (12, 0, 32, 4, "imm", False), (13, 0, 64, 2, "imm+w", False), (14, 0, 64, 2, "imm+g1500", True),
             (15, 2, 16, 5, "imm", False), (16, 2, 16, 4, "imm", False), (17, 2, 32, 3, "imm", False), (18, 3, 16, 4, "imm", False),
             (25, -3, 32, 4, "imm", False), (26, -4, 32, 4, "imm+ch2+r2", False), (27, -3, 32, 4, "imm+r2", False), (28, -3, 32, 4, "imm+r3", False),
             (29, -3, 32, 3, "imm+r4", False),
             (31, -5, 64, 2, "imm", False), (32, -5, 64, 3, "imm", False), (33, -5, 32, 4, "imm", False), (34, -5, 128, 1, "imm", False),
             (35, -5, 64, 3, "imm+w", False), (36, -5, 64, 4, "imm", False), (37, -5, 32, 6, "imm", False), (38, -5, 32, 8, "imm", False)],
    (4, 0): [(25, -3, 32, 8, "imm", False), (31, -3, 32, 16, "imm+ip", False), (32, -3, 32, 12, "imm+ip", False)],
    (4, 1): [(25, -3, 32, 8, "imm", False), (31, -3, 32, 12, "imm+ip", False), (32, -3, 32, 10, "imm+ip", False)],
    (6, 0): [(25, -3, 32, 8, "imm", False), (31, -3, 32, 16, "imm+ip", False), (32, -3, 32, 12, "imm+ip", False), (33, -3, 32, 14, "imm+ip", False)],
    (6, 1): [(25, -3, 32, 6, "imm", False), (31, -5, 64, 3, "imm", False), (32, -5, 128, 1, "imm", False), (33, -5, 64, 4, "imm", False)],
    (10, 0): [(25, -3, 32, 5, "imm", False), (26, -4, 32, 4, "imm+ch2+r2", False), (31, -3, 32, 9, "imm+ip", False), (32, -3, 32, 8, "imm+ip", False)],
    (10, 1): [(31, -5, 64, 2, "imm", False), (32, -5, 32, 3, "imm", False), (25, -3, 32, 3, "imm", False), (26, -4, 32, 4, "imm+ch2+r3", False), (27, -4, 32, 4, "imm+ch2+r4", False), (28, -4, 32, 4, "imm+ch2+r2", False)],
    (12, 0): [(31, -3, 32, 7, "imm+ip", False), (32, -3, 32, 6, "imm+ip", False), (26, -4, 32, 4, "imm+ch2+r2", False), (27, -4, 32, 4, "imm+ch2+r3", False), (2, 2, 16, 4, "imm", False), (3, 4, 16, 4, "imm", False),
              (1, 1, 16, 4, "imm", False), (20, -1, 64, 1, "imm+w", True), (30, -2, 32, 3, "imm+ch2+r4", False)],
    (12, 1): [(26, -4, 32, 3, "imm+ch2+r4", False), (27, -4, 32, 3, "imm+ch2+r2", False), (28, -4, 32, 4, "imm+ch2+r3", False),
              (32, -4, 32, 5, "imm+ch2+r3", False), (33, -4, 32, 4, "imm+ch2+r4", False), (25, -3, 32, 2, "imm", False), (2, 4, 16, 4, "imm", False), (3, 3, 16, 5, "imm", False), (4, 6, 16, 3, "imm", False), (5, 4, 32, 2, "imm", False),
              (10, 0, 32, 2, "imm", True), (20, -1, 80, 1, "imm", True), (30, -2, 16, 5, "imm+ch2+r2", False), (31, -2, 32, 2, "imm+ch4+r4", False)],
    (20, 0): [(26, -4, 32, 3, "imm+ch2+r6", False), (27, -4, 32, 3, "imm+ch2+r4", False), (2, 5, 16, 3, "imm", False), (3, 4, 16, 3, "imm", False), (4, 7, 16, 2, "imm", False),
              (1, 2, 16, 3, "imm", False), (20, -1, 48, 1, "imm", True), (30, -2, 16, 3, "imm+ch4+r6", False)],
    (20, 1): [(26, -4, 32, 2, "imm+ch2+r6", False), (27, -4, 32, 2, "imm+ch2+r4", False), (28, -4, 32, 2, "imm+ch2+r8", False), (2, 8, 16, 2, "imm", False), (3, 6, 16, 2, "imm", False), (4, 4, 16, 2, "imm", False), (5, 10, 16, 2, "imm", False),
              (1, 2, 16, 2, "imm", False), (20, -1, 32, 1, "imm", True), (30, -2, 16, 2, "imm+ch4+r6", False), (31, -2, 16, 2, "imm+ch2+r8", False)],
}
# variants under test (SFB_EXP_VARIANTS=1)
EXP = {
    # windowed loop kernel: (variant, -6, one-warp tiles per CTA, register cap)
    (8, 1): [(60, -6, 1, 168, "", False), (61, -6, 1, 255, "", False), (62, -6, 1, 200, "", False), (63, -6, 1, 232, "", False),
             (70, -6, 1, 168, "+split", False), (71, -6, 1, 128, "+split", False), (72, -6, 1, 144, "+split", False), (73, -6, 1, 200, "+split", False),
             (80, -6, 1, 255, "+rowgen", False), (81, -6, 1, 200, "+rowgen", False)],
    (8, 0): [(60, -6, 1, 168, "", False), (61, -6, 1, 128, "", False), (70, -6, 1, 128, "+split", False), (71, -6, 1, 96, "+split", False),
             (80, -6, 1, 128, "+rowgen", False)],
    (12, 1): [(70, -6, 1, 168, "+split", False), (71, -6, 1, 200, "+split", False),
              (80, -6, 1, 255, "+rowgen", False), (81, -6, 1, 200, "+rowgen", False), (82, -6, 1, 168, "+rowgen", False)],
    (12, 0): [(80, -6, 1, 128, "+rowgen", False), (81, -6, 1, 96, "+rowgen", False)],
    # round 2, measured again on the current kernels and dropped (profiles/r02_cbank.txt): "cbank" constants (LDCU.128 pairs instead
    # of UMOV immediates) for the L = 8 defaults: RK4 LROT 0.56 -> 0.74 ms, Euler 0.26 -> 0.33 ms, DDRX two-lane 0.706 -> 0.708-0.73 ms
    # 14 / 16 warps per SM for the L = 8 RK4 default (128 registers instead of 168: 450-650 bytes of spills, static FP64 share 0.71 -> 0.61):
    # 0.559 -> 0.82-0.89 ms, reduced-form I/O 0.451 -> 0.60-0.66 ms (profiles/r02_occupancy.txt)
    # tried and dropped in this session (profiles/r01_variants_sweep_a32.txt): "+ch4" loop kernels (L = 12, 20: 25-50 % slower),
    # one-lane straight-line DDRX kernels with 2-8 tiles per CTA at L = 8 (1.12 ms vs 0.73 ms for the two-lane form),
    # lock-stepped tiles "+ls" (no gain over free-running tiles that start together), two-lane L = 8 DDRX kernel with 96-node
    # tiles or per-block lock step (0.73-0.75 ms, same as the default); "+nofb" (no general-state fallback call in the kernel:
    # no spills, L = 8 RK4 0.555 -> 0.531 ms) -- the measure of what moving the fallback to a second launch could gain;
    # L2 bulk prefetch of the tile K CTAs further down the grid (K = 100..888: 0.555 -> 0.555..0.617 ms); the symmetry test folded
    # into stage 0 (mirror rows loaded in place of the redundant stage-0 n0 loads: 0.555 -> 0.729 ms, they come from DRAM inside
    # short m blocks); without any test the kernel would run at 0.485 ms (profiles/r01_notes.md); global stores in a rolled
    # epilogue loop instead of the unrolled stage body ("+ep": body 8 % shorter, L = 8 RK4 0.558 -> 0.573 ms, L = 10 0.949 -> 0.893 ms);
    # mirror rows prefetched to L2 at the head of the tile and tested in the LAST stage next to the n0 re-reads, stores in an epilogue
    # ("+lsym": 0.557 -> 0.663 ms); general tiles handed to a second launch through a work list ("+fb2": no in-kernel fallback
    # call, no spills: RK4 0.556 -> 0.537 ms, but Euler 0.260 -> 0.296 ms and general states 0.35 -> 0.56 ms per 2e5 nodes)
}
# default variant (0), chosen from the sweeps in profiles/r01_variants_sweep*.txt.
#   L <= 8 (code fits the instruction cache): straight-line kernels, small tiles, several independent CTAs per SM;
#   DDRX kernels up to L = 8 prefer four lanes per node;  L >= 12 with DDRX and every L >= 14: the table-driven loop kernel
#   (straight-line code of 150 KB .. 1 MB per stage is instruction-fetch bound, profiles/r01_notes.md).
#   R = -3: reduced one-lane kernel (real-ODF symmetry detected per tile, in-kernel two-lane fallback for general complex
#   states): half the arithmetic and shared memory per node; 1.2-1.9x the node rate wherever the straight-line form is used.
TUNE = {
    (4, 0): (-3, 32, 8, "imm", False), (4, 1): (-3, 32, 6, "imm+ip+a32+cw2", False),
    (6, 0): (-3, 32, 8, "imm+a32", False), (6, 1): (-3, 32, 3, "imm+a32+cw2", False),
    (8, 0): (-3, 32, 7, "imm+a32", False), (8, 1): (-5, 64, 3, "imm+a32", False),
    (10, 0): (-3, 32, 5, "imm+a32", False), (10, 1): (-4, 32, 4, "imm+ch2+r3", False),
    (12, 0): (-3, 32, 4, "imm+a32", False), (12, 1): (-4, 32, 4, "imm+ch2+r3", False),
    (14, 0): (-4, 32, 4, "imm+ch2+r4", False), (14, 1): (-4, 32, 3, "imm+ch2+r4", False),
    (16, 0): (-4, 32, 4, "imm+ch2+r4", False), (16, 1): (-4, 32, 3, "imm+ch2+r6", False),
    (18, 0): (-4, 32, 3, "imm+ch2+r4", False), (18, 1): (-4, 32, 2, "imm+ch2+r6", False),
    (20, 0): (-4, 32, 3, "imm+ch2+r4", False), (20, 1): (-4, 32, 2, "imm+ch2+r6", False),
}
# flags of the reduced one-lane kernel (sweep: profiles/r01_variants_sweep_a32.txt):
#   "+a32": 32-bit row strides (one multiply-add per global row address) and explicit LDG/STG in the row finalisation;
#   "+n0" (Horner RK4): n0 re-read from global in every stage -- no selects, zero-initialisation or branches around the loads;
#   "+cwN": N independent one-warp tiles per CTA (they start together and share the instruction stream's cache misses);
#   "+ip": in-place RK4 stages (one stage buffer + register delay queue).


# scheme-specific default: variant 100 (when present) replaces variant 0 for multi-stage (RK4) steps -- with DDRX the
# classical RK4 keeps three state buffers, which favours the reduced kernel's halved footprint
TUNE_RK = {
    (6, 1): (-5, 64, 3, "imm", False), (8, 1): (-5, 128, 1, "imm+a32", False),
    # LROT kernels, RK4: in-place stage update (one stage buffer + a register delay queue) -> 12 instead of 7 warps per SM at L = 8
    (4, 0): (-3, 32, 12, "imm+ip+a32+n0", False), (4, 1): (-3, 32, 8, "imm+a32", False), (6, 0): (-3, 32, 8, "imm+ip+a32+n0+cw2", False),
    (8, 0): (-3, 32, 6, "imm+ip+a32+n0+cw2", False), (10, 0): (-3, 32, 4, "imm+ip+a32+n0+cw2", False), (12, 0): (-3, 32, 6, "imm+ip+a32+n0", False),
}
# the previous full-form defaults stay selectable (variant 40) for comparisons
FULL_DEFAULT = {
    (4, 0): (1, 16, 8, "imm", False), (4, 1): (1, 16, 8, "imm", False),
    (6, 0): (1, 16, 8, "imm", False), (6, 1): (0, 32, 4, "imm", True),
    (8, 0): (1, 16, 6, "imm", False), (8, 1): (0, 64, 2, "imm", False), (10, 0): (0, 16, 4, "imm+w", True), (12, 0): (0, 16, 4, "imm+w", True),
    (10, 1): (-2, 32, 2, "imm+ch2+r4", False), (12, 1): (-2, 32, 2, "imm+ch2+r4", False),
    (14, 0): (-2, 16, 4, "imm+ch2+r4", False), (14, 1): (-2, 16, 3, "imm+ch2+r4", False),
    (16, 0): (-2, 16, 4, "imm+ch2+r6", False), (16, 1): (-2, 16, 3, "imm+ch2+r6", False),
    (18, 0): (-2, 16, 3, "imm+ch2+r6", False), (18, 1): (-2, 16, 2, "imm+ch2+r6", False),
    (20, 0): (-2, 16, 3, "imm+ch2+r6", False), (20, 1): (-2, 16, 2, "imm+ch2+r6", False),
}


def _write_if_changed(path, text):
    if os.path.exists(path) and open(path).read() == text:
        return False
    with open(path, "w") as f:
        f.write(text)
    return True


def generate(Ls):
    sys.path.insert(0, os.path.join(HERE, ".."))
    from specfab_b200.codegen import emit_step, emit_tables
    os.makedirs(GEN, exist_ok=True)
    units, metas = [], []
    for L in Ls:
        for dd in (0, 1):
            R, TN, MINB, cm0, sy0 = TUNE[(L, dd)]
            variants = [(0, R, TN, MINB, cm0, sy0)]
            if (L, dd) in TUNE_RK:
                variants.append((100,) + TUNE_RK[(L, dd)])
            if (L, dd) in FULL_DEFAULT:
                variants.append((40,) + FULL_DEFAULT[(L, dd)])
            # the tuning variants of the sweeps in profiles/ are opt-in (SFB_EXTRA_VARIANTS=1): ~100 more translation units
            if os.environ.get("SFB_EXTRA_VARIANTS", "0") == "1":
                variants += [v for v in EXTRA.get((L, dd), []) if v[1:] != variants[0][1:]]
            if os.environ.get("SFB_EXP_VARIANTS", "0") == "1":     # variants under test in the current tuning session
                variants += EXP.get((L, dd), [])
            for (vid, R, TN, MINB, cmode, sync) in variants:
                tag = "L%d_%s" % (L, "ddrx" if dd else "lrot") + ("_v%d" % vid if vid else "")
                # const_mode string: "imm" | "cbank", optional flags "+w" (register window), "+cN" (>= N DFMA
                # chains per lane), "+gN" (lock-step barrier every >= N DFMAs)
                parts = cmode.split("+")
                cm = parts[0]
                window = "w" in parts[1:]
                mc = max([int(x[1:]) for x in parts[1:] if x.startswith("c") and not x.startswith(("ch", "cw"))] + [1])
                gd = max([int(x[1:]) for x in parts[1:] if x.startswith("g")] + [0])
                if R == -6:      # windowed loop kernel (one lane per node, register window, constant-cache table); TN field = one-warp tiles per CTA, MINB field = register cap
                    from specfab_b200.codegen import emit_wloop
                    rowgen = "rowgen" in cmode.split("+")
                    wtab, wmeta = (emit_wloop.emit_rows if rowgen else emit_wloop.emit)(L, dd)
                    wname = "wtab%s_L%d_%s.inc" % ("r" if rowgen else "", L, "ddrx" if dd else "lrot")
                    _write_if_changed(os.path.join(GEN, wname), wtab)
                    meta = dict(L=L, ddrx=dd, R=1, TN=32, reduced=1, wloop=1, WPC=TN,
                                dfma_node=sum(p.dfma for p in emit_step.plan(L, dd)[1]),
                                dfma_executed=wmeta["dfma_padded"] + 8 * emit_step.nrow_phys(L) // 2, nconst=wmeta["nconst"])
                    meta["dfma_node_full"] = 2 * meta["dfma_node"]
                    split = "split" in cmode.split("+")
                    meta["split"] = int(split)
                    npl = max([int(x[3:]) for x in cmode.split("+") if x.startswith("npl")] + [1])
                    cu = ('#define SFB_ROWGEN %d\n' % int(rowgen)) + ('#define SFB_NPL %d\n' % npl) + ('#define SFB_SPLIT %d\n' % int(split)) + ('#define SFB_L %d\n#define SFB_DDRX %d\n#define SFB_WPC %d\n#define SFB_MAXREG %d\n'
                          '#define SFB_NAME sfb_launch_step_%s\n#define SFB_WTAB_INC "gen/%s"\n'
                          '#include "sfb_step_wloop.cuh"\n' % (L, dd, TN, MINB, tag, wname))
                    path = os.path.join(GEN, "step_%s.cu" % tag)
                    _write_if_changed(path, cu)
                    units.append(path)
                    meta["tag"] = tag
                    meta["variant"] = vid
                    metas.append(meta)
                    continue
                if R == -2:      # table-driven loop kernel (two lanes per node); MINB field = CTAs/SM, "chN" = rows per chunk
                    ch = max([int(x[2:]) for x in parts[1:] if x.startswith("ch")] + [2])
                    tabsrc, meta = emit_step.emit_loop_table(L, dd, ch)
                    meta.update(TN=TN, dfma_role=[0], dfma_node=2 * sum(p.dfma for p in emit_step.plan(L, dd)[1]), loads_node=0, nrow=emit_step.nrow_phys(L))
                    body = "// loop kernel: no generated body\n"
                    tab = "#define SFB_LOOP 1\n#define SFB_CH %d\n" % ch + tabsrc
                    skeleton = "sfb_step_kernel.cuh"
                    R = max([int(x[1:]) for x in parts[1:] if x.startswith("r")] + [1])     # warp roles per node group
                    meta["R"] = R
                elif R == -4:      # reduced table-driven loop kernel (one lane per node, +rN roles, +chN rows/chunk), loop-form fallback
                    ch = max([int(x[2:]) for x in parts[1:] if x.startswith("ch")] + [2])
                    tabsrc, meta = emit_step.emit_loop_table(L, dd, ch)
                    meta.update(TN=TN, dfma_role=[0], dfma_node=sum(p.dfma for p in emit_step.plan(L, dd)[1]), loads_node=0, nrow=emit_step.nrow_phys(L))
                    meta["dfma_node_full"] = 2 * meta["dfma_node"]
                    meta["reduced"] = 1
                    body = "// loop kernel: no generated body\n"
                    R = max([int(x[1:]) for x in parts[1:] if x.startswith("r")] + [1])
                    meta["R"] = R
                    tab = "#define SFB_LOOP 1\n#define SFB_CH %d\n#define SFB_REDUCED 1\n#define SFB_TNR %d\n" % (ch, TN) + tabsrc
                    skeleton = "sfb_step_kernel.cuh"
                    R_cu, TN_cu, inc_cu = R, 16, "gen/apply_%s.inc" % tag
                elif R == -5:      # reduced two-lane (re|im) kernel derived from the four-lane form; four-lane fallback on TN/2 nodes
                    body, tab, meta = emit_step.emit4(L, dd, TN, cm, sync, window, mc, gd, reduced=True)
                    fbody, ftab, fmeta = emit_step.emit4(L, dd, TN // 2, cm, sync, window, mc, gd)
                    _write_if_changed(os.path.join(GEN, "apply_%s_full.inc" % tag), fbody)
                    tab = ('#define SFB_REDUCED 1\n#define SFB_TNR %d\n#define SFB_APPLY_INC_R "gen/apply_%s.inc"\n' % (TN, tag)) + tab
                    if "a32" in parts[1:]:
                        tab = "#define SFB_A32 1\n" + tab
                    skeleton = "sfb_step_kernel4.cuh"
                    meta["dfma_node_full"] = fmeta["dfma_node"]
                    meta["reduced"] = 1
                    meta["R"] = 1
                    R_cu, TN_cu, inc_cu = 0, TN // 2, "gen/apply_%s_full.inc" % tag
                elif R == -3:      # reduced one-lane kernel for real-ODF states (+ in-kernel two-lane fallback, tiles of 16)
                    Rr = max([int(x[1:]) for x in parts[1:] if x.startswith("r")] + [1])      # "+rN": warp roles sharing the 32 nodes
                    ip = "ip" in parts[1:]             # "+ip": in-place stage update (one stage buffer; single role)
                    ls = "ls" in parts[1:]             # "+ls" (with "+cwN"): the CTA's tiles run the reduced body in lock step
                    body, tab, meta = emit_step.emit(L, dd, Rr, TN, cm, ls, mc, gd, reduced=True, inplace=ip)
                    if ls:
                        tab = "#define SFB_LS 1\n" + tab
                    fbody, _, fmeta = emit_step.emit(L, dd, Rr, 16, cm, False, mc, gd, inplace=ip)
                    if ip:
                        tab = "#define SFB_INPLACE 1\n" + tab
                    cw = max([int(x[2:]) for x in parts[1:] if x.startswith("cw")] + [1])
                    if cw > 1:                     # "+cwN": N independent one-warp tiles per CTA (MINB then counts CTAs of N warps)
                        tab = "#define SFB_CW %d\n" % cw + tab
                    if "nofb" in parts[1:]:        # experiment: no general-state fallback (traps)
                        tab = "#define SFB_NOFB 1\n" + tab
                    if "n0" in parts[1:]:          # "+n0": n0 loaded from global in every stage (Horner kernels), no selects
                        tab = "#define SFB_N0ALL_REQ 1\n" + tab
                    if "a32" in parts[1:]:         # "+a32": 32-bit row strides + explicit global loads/stores in the row finalisation
                        tab = "#define SFB_A32 1\n" + tab
                    _write_if_changed(os.path.join(GEN, "apply_%s_full.inc" % tag), fbody)
                    tab = ('#define SFB_REDUCED 1\n#define SFB_TNR %d\n#define SFB_APPLY_INC_R "gen/apply_%s.inc"\n' % (TN, tag)) + tab
                    skeleton = "sfb_step_kernel.cuh"
                    meta["dfma_node_full"] = fmeta["dfma_node"]
                    meta["reduced"] = 1
                    meta["R"] = Rr
                    R_cu, TN_cu, inc_cu = Rr, 16, "gen/apply_%s_full.inc" % tag
                elif R == -1:      # persistent, lock-stepped, streaming refill (four lanes per node)
                    body, tab, meta = emit_step.emit4(L, dd, TN, cm, True, window, mc, gd)
                    skeleton = "sfb_step_kernel5.cuh"
                    tab = "#define SFB_WINDOW %d\n" % int(window) + tab
                elif R == 0:
                    body, tab, meta = emit_step.emit4(L, dd, TN, cm, sync, window, mc, gd)
                    skeleton = "sfb_step_kernel4.cuh"
                else:
                    body, tab, meta = emit_step.emit(L, dd, R, TN, cm, sync, mc, gd)
                    skeleton = "sfb_step_kernel.cuh"
                _write_if_changed(os.path.join(GEN, "apply_%s.inc" % tag), body)
                if meta.get("reduced"):
                    r_cu, tn_cu, inc = R_cu, TN_cu, inc_cu
                else:
                    r_cu, tn_cu, inc = R, TN, "gen/apply_%s.inc" % tag
                cu = ('#define SFB_L %d\n#define SFB_DDRX %d\n#define SFB_R %d\n#define SFB_TN %d\n#define SFB_MINB %d\n'
                      '#define SFB_NAME sfb_launch_step_%s\n#define SFB_APPLY_INC "%s"\n%s'
                      '#include "%s"\n' % (L, dd, r_cu, tn_cu, MINB, tag, inc, tab, skeleton))
                path = os.path.join(GEN, "step_%s.cu" % tag)
                _write_if_changed(path, cu)
                units.append(path)
                meta["tag"] = tag
                meta["variant"] = vid
                metas.append(meta)
    reg = ["// GENERATED registry of step launchers"]
    for m in metas:
        reg.append('extern "C" cudaError_t sfb_launch_step_%s(const SfbStepParams&, const SfbRegConst&, cudaStream_t);' % m["tag"])
    reg.append("static const SfbStepEntry kStepRegistry[] = {")
    for m in metas:
        reg.append("  {%d, %d, %d, %d, %d, %d, sfb_launch_step_%s}," % (m["L"], m["ddrx"], m["variant"], m["R"], m["TN"], m["dfma_node"], m["tag"]))
    reg.append("};")
    _write_if_changed(os.path.join(GEN, "registry.inc"), "\n".join(reg) + "\n")
    _write_if_changed(os.path.join(GEN, "tables.inc"), emit_tables.emit())
    _write_if_changed(os.path.join(GEN, "orth_tables.inc"), emit_tables.emit_orthotropic())
    _write_if_changed(os.path.join(GEN, "moments_hi.inc"), emit_tables.emit_moments_hi())
    _write_if_changed(os.path.join(GEN, "ingest.inc"), emit_tables.emit_ingest())
    with open(os.path.join(GEN, "meta.json"), "w") as f:
        json.dump(metas, f, indent=1)
    return units, metas


def _deps_hash(src):
    h = hashlib.sha1()
    files = [src] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files += [os.path.join(HERE, "..", "include", "specfab_b200.h")]
    if "step_" in os.path.basename(src):
        tag = os.path.basename(src)[5:-3]
        if os.path.exists(os.path.join(GEN, "apply_%s.inc" % tag)):
            files.append(os.path.join(GEN, "apply_%s.inc" % tag))
        base = tag.split("_v")[0]
        for wt in ("wtab_%s.inc" % base, "wtabr_%s.inc" % base):
            if os.path.exists(os.path.join(GEN, wt)):
                files.append(os.path.join(GEN, wt))
        if os.path.exists(os.path.join(GEN, "apply_%s_full.inc" % tag)):
            files.append(os.path.join(GEN, "apply_%s_full.inc" % tag))
    else:
        files += [os.path.join(GEN, "registry.inc"), os.path.join(GEN, "tables.inc"), os.path.join(GEN, "orth_tables.inc"), os.path.join(GEN, "moments_hi.inc"), os.path.join(GEN, "ingest.inc")]
    for f in files:
        h.update(open(f, "rb").read())
    h.update(" ".join(CFLAGS + ARCH).encode())
    return h.hexdigest()


def compile_one(src, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".sha1"
    hsh = _deps_hash(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == hsh:
        return obj, 0.0, ""
    t0 = time.time()
    cmd = [NVCC] + ARCH + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    with open(stamp, "w") as f:
        f.write(hsh)
    return obj, time.time() - t0, p.stderr


def build(Ls=None, jobs=None, verbose=False):
    Ls = Ls or ALL_L
    jobs = jobs or max(1, (os.cpu_count() or 2))
    units, metas = generate(Ls)
    units = units + [os.path.join(CSRC, "sfb_api.cu"), os.path.join(CSRC, "sfb_fields.cu"), os.path.join(CSRC, "sfb_operators.cu"),
                     os.path.join(CSRC, "sfb_orthotropic.cu"), os.path.join(CSRC, "sfb_fields_hi.cu")]
    objs = []
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for obj, dt, log in ex.map(lambda s: compile_one(s, verbose), units):
            objs.append(obj)
            if dt:
                print("  compiled %-28s %6.1fs" % (os.path.basename(obj), dt), flush=True)
            if verbose and log:
                print(log)
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    print("linked", LIB)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=None)
    ap.add_argument("--L", type=str, default=None, help="comma list of truncations to build (default: all even 4..20)")
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    build([int(x) for x in a.L.split(",")] if a.L else None, a.j, a.v)
