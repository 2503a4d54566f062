"""CPU: libfm_b200.so builds for sm_100a, loads, and exports every symbol that
include/fm_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from planning_motion_planning_b200 import _capi, build


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fmb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("fmb_solve2d_f64", "fmb_solve2d_f32", "fmb_solve3d_f64", "fmb_trace2d_f64", "fmb_trace3d_f64",
              "fmb_finish", "fmb_last_error", "fmb_version"):
        assert s in syms


def test_library_builds_and_exports_all_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in fm_b200.h but not exported"
    # the binding covers exactly the declared ABI
    assert sorted(_capi.SIGNATURES) == declared_symbols()
    assert lib.fmb_version() >= 100


def test_library_is_sm100a_sass():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_validation_without_gpu():
    lib = _capi.lib()
    assert lib.fmb_workspace_bytes_2d(0, 10, 1) == 0
    assert lib.fmb_workspace_bytes_2d(4096, 4096, 1) > 4096 * 4096 // (16 * 32) * 4
    rc = lib.fmb_solve2d_f64(None, 0, 0, None, 0, 0, 10, 10, 1, None, None, 0, None)
    assert rc == _capi.FMB_E_INVALID and b"null" in lib.fmb_last_error()
