"""The C-ABI shared library loads and exports every symbol include/dynam3d_b200.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dynam3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(d3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dynam3d_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert getattr(lib, n, None) is not None, f"{n} declared in the header but not exported"
    assert lib.d3d_version() >= 100
    assert isinstance(lib.d3d_last_error(), bytes)


def test_python_signatures_cover_the_header():
    from dynam3d_b200 import _lib
    declared = set(_declared()) - {"d3d_last_error", "d3d_gemm_args"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_oracle_import_in_product_package():
    """The product path must never route through the oracle (checker only)."""
    pkg = os.path.join(ROOT, "dynam3d_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_missing_library_fails_loudly(monkeypatch):
    from dynam3d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdynam3d_b200.so")
    try:
        _lib.lib()
        assert False, "expected D3DLibraryError"
    except _lib.D3DLibraryError:
        pass
