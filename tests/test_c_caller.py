"""A plain-C program (tests/c/abi_smoke.c, gcc -std=c99 -pedantic) against include/specfab_b200.h + libspecfab_b200.so: the
header is valid C, the library links without Python or torch, reports the missing device instead of falling back (CPU), and
gives the same bits as the Python binding (GPU)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_c(tmp_path):
    from specfab_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lspecfab_b200", "-Wl,-rpath," + libdir]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def test_header_is_c99_and_library_fails_loudly_without_a_device(tmp_path):
    exe = build_c(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    p = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.startswith("init -3") and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_c_caller_matches_python_binding(tmp_path):
    import specfab_b200 as sf
    exe = build_c(tmp_path)
    N = 37
    p = subprocess.run([exe, str(N)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.strip().splitlines()
    assert lines[0] == "init 0" and lines[1] == "step 0"
    vals = np.array([[float.fromhex(t) for t in l.split()] for l in lines[2:]])
    lm, n = sf.init(8)
    assert vals.shape == (n, 2)
    x = np.zeros((N, n), dtype=np.complex128)
    x[:, 0] = 0.28209479177387814
    ug = np.zeros((N, 3, 3))
    ug[:] = np.diag([0.5, 0.5, -1.0])
    y = sf.step_arr(x, ug, dt=0.01, terms=("lrot", "reg"), nsteps=5)
    assert np.array_equal(vals[:, 0] + 1j * vals[:, 1], y[0])


@pytest.mark.gpu
def test_c_caller_multi_device_entry_point(tmp_path):
    """sfb_step_arr_multi from plain C over every device of the box: same bits as the one-device call."""
    exe = build_c(tmp_path)
    N = 1000
    a = subprocess.run([exe, str(N)], capture_output=True, text=True, timeout=300)
    b = subprocess.run([exe, str(N), "multi"], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, a.stdout + b.stdout + b.stderr
    assert a.stdout == b.stdout and a.stdout.splitlines()[1] == "step 0"
