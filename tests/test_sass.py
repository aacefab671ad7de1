"""CPU (nvcc cross-compiles): SASS-level pins.

1. The windowed loop kernel (csrc/sfb_step_wloop.cuh, experimental variants) must read its operator table through the
   UNIFORM datapath: LDCU c[3][UR + imm] feeding DFMA UR operands.  ptxas 12.9 silently falls back to per-lane LDC -- a third
   of the speed -- unless the warp is provably converged in front of the loop (__syncwarp) and the table position / warp role
   come out of a warp reduction (REDUX -> uniform register); profiles/r02_notes.md has the bisection.  This test keeps that
   finding executable.
2. The default step kernels stage their tiles with 1-D TMA bulk copies (UBLKCP) and wait on mbarriers (SYNCS), as DESIGN.md
   section 2 claims (profiles/r02_sass_digest.json is the committed digest of the same objects)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "specfab_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def loops_of(sass, kernel):
    """[(n_instr, counter)] of every backward-branch loop body longer than 256 instructions in `kernel`"""
    import collections
    fn = [f for f in re.split(r"Function : ", sass)[1:] if kernel in f.split("\n")[0]][0]
    ins = []
    for ln in fn.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    out = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > 256 * 16:
            lo = int(m.group(1), 16)
            c = collections.Counter()
            for b, u in ins:
                if lo <= b <= a:
                    u = re.sub(r"^@!?U?P\d+\s+", "", u)
                    c[u.split()[0].split(".")[0]] += 1
            out.append(((a - lo) // 16, c))
    return out


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
@pytest.mark.parametrize("split,maxreg", [(0, 255), (1, 168)])
def test_windowed_loop_reads_its_table_through_the_uniform_datapath(tmp_path, split, maxreg):
    sys.path.insert(0, ROOT)
    from specfab_b200.codegen import emit_wloop
    tab, meta = emit_wloop.emit(8, 1)
    (tmp_path / "wtab.inc").write_text(tab)
    cu = tmp_path / "w.cu"
    cu.write_text('#define SFB_SPLIT %d\n#define SFB_L 8\n#define SFB_DDRX 1\n#define SFB_WPC 1\n#define SFB_MAXREG %d\n'
                  '#define SFB_NAME sfb_launch_step_test\n#define SFB_WTAB_INC "%s"\n#include "sfb_step_wloop.cuh"\n'
                  % (split, maxreg, str(tmp_path / "wtab.inc")))
    cubin = str(tmp_path / "w.cubin")
    p = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-I", os.path.join(ROOT, "include"),
                        "-cubin", "-o", cubin, str(cu)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    kernel = "step_kernel_w2rE" if split else "step_kernel_wrE"
    bodies = [c for n, c in loops_of(sass, kernel) if c["DFMA"] >= 400]
    assert bodies, "no operator loop found"
    for c in bodies:
        assert c["LDCU"] >= 100 and c["LDC"] <= c["LDCU"] // 10, dict(LDCU=c["LDCU"], LDC=c["LDC"], DFMA=c["DFMA"])
    assert meta["dfma_padded"] == 5260 and meta["dfma_useful"] == 4316 and meta["nconst"] == 1605


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
def test_row_generic_form_has_one_small_body_for_every_L(tmp_path):
    """SFB_ROWGEN at L = 12: the operator loop is ONE row (79 table entries through LDCU, ~240 DFMA) whatever L is."""
    sys.path.insert(0, ROOT)
    from specfab_b200.codegen import emit_wloop
    tab, meta = emit_wloop.emit_rows(12, 1)
    assert meta["stride"] == 79 and meta["nconst"] == 79 * 49 and meta["dfma_useful"] == 9374
    (tmp_path / "wtab.inc").write_text(tab)
    cu = tmp_path / "w.cu"
    cu.write_text('#define SFB_ROWGEN 1\n#define SFB_L 12\n#define SFB_DDRX 1\n#define SFB_WPC 1\n#define SFB_MAXREG 255\n'
                  '#define SFB_NAME sfb_launch_step_test\n#define SFB_WTAB_INC "%s"\n#include "sfb_step_wloop.cuh"\n' % str(tmp_path / "wtab.inc"))
    cubin = str(tmp_path / "w.cubin")
    p = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-I", os.path.join(ROOT, "include"),
                        "-cubin", "-o", cubin, str(cu)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    bodies = [(n, c) for n, c in loops_of(sass, "step_kernel_wrE") if c["DFMA"] >= 200]
    assert bodies, "no operator loop found"
    n, c = min(bodies, key=lambda b: b[0])
    assert n < 1200 and 230 <= c["DFMA"] <= 300 and c["LDCU"] >= 79 and c["LDC"] <= 12, (n, dict(c))


def test_default_step_kernels_use_tma_bulk_copies_and_mbarriers():
    obj = os.path.join(ROOT, "specfab_b200", "_build", "step_L8_lrot.o")
    if not os.path.exists(obj):
        import __graft_entry__
        __graft_entry__.build()
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass and "DFMA" in sass
    assert "HMMA" not in sass and "UTCHMMA" not in sass          # an FP64 CUDA-core path: no tensor-core instruction expected
