"""CPU tests for the boundary: the C-ABI library loads and exports every symbol that
include/qcsim_b200.h declares, with no compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qcsim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qcsim_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from qcsim_b200 import _lib, build

    build.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/qcsim_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in qcsim_b200/_lib.py"
    assert lib.qcsim_abi_version() == 2


def test_product_fails_loudly_without_gpu():
    import qcsim_b200
    from qcsim_b200 import _lib

    lib = _lib.load()
    n = C.c_int(-1)
    rc = lib.qcsim_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(qcsim_b200.QcsimError) as e:
        qcsim_b200.QubitRegister(4)
    assert e.value.code == _lib.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The product path must not route through the checker."""
    pkg = os.path.join(ROOT, "qcsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "libqcsim_oracle" not in src and "libqcsim_ref" not in src and "qcsim_oracle.c" not in src, f
