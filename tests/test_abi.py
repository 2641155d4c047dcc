"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_ \*]*?\b((?:zng_|zlibng_)[A-Za-z0-9_]+)\s*\(", re.M)


def declared_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(inc, fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in DECL.finditer(src):
            syms.add(m.group(1))
    return syms


def test_library_exports_every_declared_symbol(pkg):
    L = ctypes.CDLL(pkg.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20, syms
    missing = [s for s in sorted(syms) if not hasattr(L, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_python_binding_resolves(pkg):
    pkg.lib()
    assert pkg.deflate_bound(65536) % 16 == 0 and pkg.deflate_bound(65536) >= 65536 + 8192 + 16


def test_no_device_is_an_error_not_a_fallback(pkg):
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(pkg.ZngB200Error):
        pkg.Context(0)
