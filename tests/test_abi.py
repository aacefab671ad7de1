"""CPU: the C-ABI library loads, exports every symbol include/specfab_b200.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "specfab_b200.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    from specfab_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 25
    for nm in names:
        assert hasattr(lib, nm), "missing export " + nm
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"


def test_fails_loudly_without_cuda_or_reports_init_errors():
    import specfab_b200 as sf
    from specfab_b200 import _lib
    lib = _lib.load()
    with pytest.raises(sf.SpecfabB200Error) as ei:
        sf.init(7)                      # odd L is rejected before touching the device
    assert ei.value.code == _lib.SFB_EINVAL
    if lib.sfb_device_count() == 0:
        with pytest.raises(sf.SpecfabB200Error) as ei:
            sf.init(8)
        assert ei.value.code == _lib.SFB_ECUDA and "no CPU fallback" in str(ei.value)


def test_build_info_lists_every_truncation():
    import specfab_b200 as sf
    info = sf.build_info()
    Ls = sorted({k["L"] for k in info["step_kernels"]})
    assert Ls == [4, 6, 8, 10, 12, 14, 16, 18, 20] and info["arch"] == "sm_100a"


def test_product_does_not_import_the_oracle():
    """the product path must never route through oracle/ (parity claims would be void)"""
    pkg = os.path.join(ROOT, "specfab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and "gen" not in dirpath:
                src = open(os.path.join(dirpath, f)).read()
                assert "import specfab_oracle" not in src and "oracle_c" not in src and "liboracle" not in src, f


def test_fortran_shim_binds_only_declared_symbols():
    """every bind(c, name='...') of the iso_c_binding shim must be a symbol the header declares"""
    src = open(os.path.join(ROOT, "fortran", "specfab_b200.f90")).read()
    bound = sorted(set(re.findall(r"bind\(c,\s*name='(sfb_[A-Za-z0-9_]+)'\)", src)))
    names = declared_functions()
    assert len(bound) >= 10
    for nm in bound:
        assert nm in names, "fortran shim binds an undeclared symbol " + nm
    for nm in ("sfb_init", "sfb_step_arr", "sfb_step_rnlm_arr", "sfb_Eij_tranisotropic_arr"):
        assert nm in bound


def test_compute_entry_points_refuse_to_run_uninitialised():
    """without a device sfb_init fails, and every compute entry point must then return SFB_ENOINIT -- no crash, no CPU path"""
    import ctypes as C
    from specfab_b200 import _lib
    lib = _lib.load()
    if lib.sfb_device_count() > 0:
        pytest.skip("device present: the library may be initialised by other tests")
    assert lib.sfb_init(8) == _lib.SFB_ECUDA
    o = _lib.StepOpts()
    o.dt, o.terms, o.scheme, o.nsteps = 0.1, 1, 1, 1
    buf = (C.c_double * 4096)()
    p = C.cast(buf, C.c_void_p)
    calls = [
        lambda: lib.sfb_step_arr(p, p, 1, 1, p, None, C.byref(o)),
        lambda: lib.sfb_step_rnlm_arr(p, p, 1, 1, p, None, C.byref(o)),
        lambda: lib.sfb_step_arr_dev(p, p, 1, 1, 1, p, 1, None, 1, C.byref(o), None),
        lambda: lib.sfb_step_rnlm_arr_dev(p, p, 1, 1, 1, p, 1, None, 1, C.byref(o), None),
        lambda: lib.sfb_a2_arr(p, 1, 1, p),
        lambda: lib.sfb_a4_arr(p, 1, 1, p),
        lambda: lib.sfb_eig_arr(p, 1, 1, p, p),
        lambda: lib.sfb_Eij_eigenframe_arr(p, 1, 1, p, 0.0125, 1, p, None, None, None),
        lambda: lib.sfb_Eij_eigenframe_rnlm_arr_dev(p, 1, 1, p, 0.0125, 1, p, None, None, None, None, None),
    ]
    for f in calls:
        assert f() == _lib.SFB_ENOINIT
    assert b"sfb_init" in lib.sfb_last_error()
