"""The C-ABI library loads (no GPU needed) and exports every symbol include/kolm_abi.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kolm_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kolm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from kolmogorovlike_datacompressor_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.so_path())
    syms = declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(L, s), s
    assert L.kolm_abi_version() >= 1
    L.kolm_strerror.restype = ctypes.c_char_p
    assert L.kolm_strerror(0) == b"ok"
    assert set(_lib.EXPORTS) <= set(syms)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "kolmogorovlike_datacompressor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libkolm_oracle" not in txt, f
