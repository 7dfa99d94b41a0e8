"""C-ABI library: loads, exports every symbol include/gpbo.h declares, fails loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from gpbo_pkg import pkg


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gpbo.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpbo_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    lib = pkg._lib.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gpbo.h but not exported"
    assert set(syms) == set(pkg._lib.EXPORTS)


def test_version_and_error_string():
    lib = pkg._lib.load()
    assert lib.gpbo_version() >= 100
    assert isinstance(lib.gpbo_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.GpboError, match="no CUDA device|CUDA"):
        pkg.Context(0)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(pkg._lib, "_lib", None)
    monkeypatch.setattr(pkg._lib, "LIB_PATH", "/nonexistent/libgpbo.so")
    with pytest.raises(pkg.GpboError, match="no CPU fallback"):
        pkg._lib.load()


def test_bad_arguments_are_rejected_before_any_device_work():
    lib = pkg._lib.load()
    assert lib.gpbo_create(None, 0, 0) == -1
    assert b"NULL" in lib.gpbo_last_error()
    x = np.zeros(3)
    f = ctypes.c_double()
    rc = lib.gpbo_lbfgsb_minimize(ctypes.cast(None, pkg._lib.OBJECTIVE_FN), None, None, None, None,
                                  x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(f), None, None, None)
    assert rc == -1


def test_product_never_imports_the_oracle_or_a_cpu_solver():
    """The product path may not route through oracle/, sklearn or scipy (no CPU fallback)."""
    pkgdir = os.path.join(ROOT, "gp-bayesopinf_b200")
    bad = re.compile(r"^\s*(import|from)\s+(gp_oracle|ref_import|oracle|sklearn|scipy)\b", re.M)
    for dirpath, _, files in os.walk(pkgdir):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not bad.search(src), fn


def test_header_is_valid_c_and_cxx(tmp_path):
    """include/gpbo.h must be consumable by a plain C compiler (the ABI has no C++ or torch types)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    hdr = os.path.join(ROOT, "include", "gpbo.h")
    c = tmp_path / "t.c"
    c.write_text('#include "gpbo.h"\nint main(void) { gpbo_ctx* c = 0; return gpbo_destroy(c) + (GPBO_NCLASS > 0 ? 0 : 1); }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.dirname(hdr), str(c)],
                   check=True)
    cc = tmp_path / "t.cpp"
    cc.write_text('#include "gpbo.h"\nint main() { return GPBO_OK; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.dirname(hdr), str(cc)], check=True)
