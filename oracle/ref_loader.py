"""Test infrastructure only: import the UNMODIFIED Python reference by path.

The reference lives at /root/reference (read-only, absent on the GPU box).  Only
`tests/golden/make_golden.py` and the `-m "not gpu"` cross-validation tests (which
skip when the reference is absent) may use this loader; nothing in the product
package imports it.

    KF  = final/kolm_final.py                          ('KOLM' container)
    V22 = final_researched/kolm_final_researched_v2-2.py ('KOLR' container)
"""
from __future__ import annotations

import importlib.util
import os
import sys

REF_DIR = os.environ.get("KOLM_REF_DIR", "/root/reference")
_KF_PATH = os.path.join(REF_DIR, "final", "kolm_final.py")
_V22_PATH = os.path.join(REF_DIR, "final_researched", "kolm_final_researched_v2-2.py")
_cache = {}


def available() -> bool:
    return os.path.isfile(_KF_PATH) and os.path.isfile(_V22_PATH)


def _load(name: str, path: str):
    if name in _cache:
        return _cache[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def load_kf():
    """kolm_final.py as a module (compress/decompress, _ENCODERS/_DECODERS, stage functions)."""
    return _load("_ref_kolm_final", _KF_PATH)


def load_v22():
    """kolm_final_researched_v2-2.py as a module (file name is not importable by `import`)."""
    return _load("_ref_kolm_v22", _V22_PATH)
