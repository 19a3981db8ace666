"""CPU: the C-ABI library builds, loads and exports every symbol include/stswin_b200.h declares."""
import os
import re

from stswincl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "stswin_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stswin_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    _lib.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 3
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/stswin_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in stswincl_b200/_lib.py"
    assert lib.stswin_abi_version() >= 1
    assert lib.stswin_last_error() is not None


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "stswincl_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
