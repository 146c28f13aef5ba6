"""The C-ABI library loads, exports every symbol include/gr2m_b200.h declares, and refuses to
compute without a CUDA device (no CPU fallback).  Runs on the CPU box."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from gr2msemidistr_b200 import capi as m
    return m


def _declared():
    src = open(os.path.join(ROOT, "include", "gr2m_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:gr2m|wfa)_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(capi):
    declared = _declared()
    assert len(declared) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in gr2m_b200.h but not exported: {missing}"
    assert sorted(capi.SYMBOLS) == declared, "capi.SYMBOLS and the header drifted apart"
    lib = capi.lib()
    assert lib.gr2m_version() == 1


def test_flag_values_match_header_and_oracle(capi, oracle):
    src = open(os.path.join(ROOT, "include", "gr2m_b200.h")).read()
    defs = dict(re.findall(r"#define\s+((?:GR2M|WFA)_[A-Z0-9_]+)\s+(-?\d+)", src))
    assert int(defs["GR2M_PERC_F32_EXPONENT"]) == capi.PERC_F32_EXPONENT == oracle.PERC_F32_EXPONENT
    assert int(defs["GR2M_SINK_UNROUNDED"]) == capi.SINK_UNROUNDED == oracle.SINK_UNROUNDED
    assert int(defs["GR2M_MATH_LITERAL"]) == capi.MATH_LITERAL
    assert (int(defs["GR2M_OF_KGE"]), int(defs["GR2M_OF_NSE"]), int(defs["GR2M_OF_RMSE"])) == (0, 1, 2)
    assert int(defs["WFA_F32"]) == capi.WFA_F32 and int(defs["WFA_NO_CONTAMINATION"]) == capi.WFA_NO_CONTAMINATION
    assert int(defs["GR2M_ERR_NODEVICE"]) == capi.ERR_NODEVICE


def test_argument_validation_needs_no_gpu(capi):
    lib = capi.lib()
    h = C.c_void_p()
    one = np.ones(4)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    nd, reg = np.full(2, 30, np.int32), np.zeros(2, np.int32)
    assert lib.gr2m_basin_create(0, 2, dp(one), dp(one), dp(one), ip(nd), ip(reg), 1, C.byref(h)) == capi.ERR_INVALID
    assert b"positive" in lib.gr2m_last_error()
    bad = np.array([0, 3], np.int32)
    assert lib.gr2m_basin_create(2, 2, dp(one), dp(one), dp(one), ip(nd), ip(bad), 1, C.byref(h)) == capi.ERR_INVALID
    assert b"region_id" in lib.gr2m_last_error()
    assert lib.gr2m_basin_create(2, 2, None, dp(one), dp(one), ip(nd), ip(reg), 1, C.byref(h)) == capi.ERR_INVALID
    assert lib.wfa_plan_create(None, 4, 4, 0, C.byref(h)) == capi.ERR_INVALID
    assert lib.gr2m_run(None, dp(one), None, None, 0, None, None, None, None, None, None, None) == capi.ERR_INVALID


def test_no_cpu_fallback(capi):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.GR2MError) as ei:
        capi.Basin(np.ones((2, 3)), np.ones((2, 3)), np.ones(2), np.full(3, 30), np.zeros(2), 1)
    assert ei.value.code == capi.ERR_NODEVICE and "no CPU path" in str(ei.value)
    with pytest.raises(capi.GR2MError) as ei:
        capi.Plan(np.ones((4, 4), np.int16))
    assert ei.value.code == capi.ERR_NODEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gr2msemidistr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "gr2m_oracle" not in txt and "orc_" not in txt, f


def test_header_is_plain_c_and_a_c_caller_links(capi, tmp_path):
    """include/gr2m_b200.h must be consumable from C (R's .Call shim, cgo, ...): build examples/c_abi_demo.c
    as strict C99 against the shared library and run it; without a GPU it reports that and exits 0."""
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", libdir, "-lgr2m_b200",
                           f"-Wl,-rpath,{libdir}", "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ABI version 1" in out.stdout
    if capi.device_count() == 0:
        assert "no GPU here" in out.stdout
    else:
        assert "set 0: 1 - NSE = 0.000" in out.stdout
