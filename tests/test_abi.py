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


def test_headers_are_plain_c_and_link_against_the_library(pkg, tmp_path):
    """include/*.h are the drop-in boundary: they must compile as strict C11 (no C++, no torch types) and a C program that
    uses both surfaces must link against libzng_b200.so alone."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        import pytest
        pytest.skip("no C compiler")
    src = tmp_path / "use_headers.c"
    src.write_text(r'''
#include "zlib-ng.h"
#include "zng_b200.h"
#include <stdio.h>
int main(void) {
    zng_stream s;
    struct zng_b200_functable ft;
    (void)s; (void)ft;
    printf("%s %d %zu %zu\n", zlibng_version(), zng_b200_device_count(), zng_b200_deflate_bound(65536), sizeof(zng_stream));
    return zng_deflateInit2(NULL, 1, Z_DEFLATED, 15, 8, Z_DEFAULT_STRATEGY) == Z_STREAM_ERROR ? 0 : 1;
}
''')
    exe = tmp_path / "use_headers"
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.run([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lzng_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, check=True).stdout.decode().split()
    assert out[0].startswith("2.2.2") and int(out[2]) == pkg.deflate_bound(65536) and int(out[3]) == 104
