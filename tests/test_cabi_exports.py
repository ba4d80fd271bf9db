"""The C-ABI library loads on a machine without a GPU and exports every symbol include/gritas_b200.h
declares.  No compute entry point is called here."""
import ctypes
import os
import re

from gritas_b200 import build, cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "gritas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gritas_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = header_functions()
    assert len(names) >= 20
    raw = ctypes.CDLL(build.LIB)
    for nm in names:
        assert hasattr(raw, nm), f"{nm} declared in include/gritas_b200.h but not exported"


def test_python_binding_covers_the_header(lib):
    assert sorted(cabi.SYMBOLS) == header_functions()


def test_version_and_null_handle(lib):
    assert lib.gritas_b200_version() >= 100
    assert lib.gritas_b200_last_error(None) == b"null handle"
    assert lib.gritas_b200_launch_count(None) == 0


def test_create_argument_validation(lib):
    """Bad arguments are rejected before any CUDA call (works without a device)."""
    import numpy as np
    h = ctypes.c_void_p()
    p = np.array([2000.0, 0.0])
    t = np.zeros((4, 4), np.int32, order="F")
    rc = lib.gritas_b200_create(ctypes.byref(h), 0, 10, 1, 1, 0, 1, 1, p.ctypes.data, p.ctypes.data, t.ctypes.data, 4, 4, 0, -1)
    assert rc == cabi.ERR_ARG and not h.value
    rc = lib.gritas_b200_create(ctypes.byref(h), 8, 5, 1, 1, 0, 1, 4, p.ctypes.data, p.ctypes.data, t.ctypes.data, 4, 4, 0, -1)
    assert rc == cabi.ERR_NR        # 'could not handle more then 2 residuals' (m_gritas_grids.f:395-396)
    rc = lib.gritas_b200_create(ctypes.byref(h), 8, 5, 1, 1, 0, 1, 1, p.ctypes.data, p.ctypes.data, t.ctypes.data, 4, 4, 7, -1)
    assert rc == cabi.ERR_ARG       # unknown mode


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device create() must fail with ERR_CUDA -- never run on the CPU."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    p1 = np.array([2000.0, 500.0])
    p2 = np.array([500.0, 0.0])
    t = np.zeros((4, 4), np.int32, order="F")
    rc = lib.gritas_b200_create(ctypes.byref(h), 8, 5, 2, 2, 0, 1, 1, p1.ctypes.data, p2.ctypes.data, t.ctypes.data, 4, 4, 0, -1)
    assert rc == cabi.ERR_CUDA and not h.value
    # the pre-sort works without a handle, but not without a device: it must not sort on the host
    n = 16
    ic = np.arange(n, dtype=np.int32)
    dc = np.linspace(-1.0, 1.0, n)
    out = np.full(n, -7, np.int32)
    rc = lib.gritas_b200_presort(None, n, ic.ctypes.data, dc.ctypes.data, dc.ctypes.data, dc.ctypes.data, ic.ctypes.data,
                                 ic.ctypes.data, ic.ctypes.data, ic.ctypes.data, 0, out.ctypes.data)
    assert rc == cabi.ERR_CUDA and np.all(out == -7)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under gritas_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "gritas_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                pat = r"^\s*(from\s+\S*oracle|import\s+\S*oracle|#include.*oracle)|(dlopen|CDLL)\(.*oracle"
                assert not re.search(pat, src, flags=re.M), f"{f} references oracle/"
