"""Host-side logic of the C11 library that needs no GPU: the combine algebra (crc32_braid_comb.c:16-24, adler32.c:32-54), the
bounds, the argument checks -- and that, without a device, the compute entry points fail loudly instead of falling back."""
import ctypes
import zlib as pyzlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L(pkg):
    return pkg.lib()


def test_crc32_combine_family(pkg, L):
    rng = np.random.default_rng(1)
    for la, lb in ((0, 0), (1, 0), (0, 1), (1, 1), (5, 65536), (65536, 65536), (12345, 7), (1 << 20, 3), (777, 1 << 22)):
        a = rng.integers(0, 256, size=la, dtype=np.uint8).tobytes(); b = rng.integers(0, 256, size=lb, dtype=np.uint8).tobytes()
        ca, cb = pyzlib.crc32(a), pyzlib.crc32(b)
        assert L.zng_crc32_combine(ca, cb, lb) == pyzlib.crc32(a + b), (la, lb)
        assert L.zng_crc32_combine_op(ca, cb, L.zng_crc32_combine_gen(lb)) == pyzlib.crc32(a + b), (la, lb)
        aa, ab = pyzlib.adler32(a), pyzlib.adler32(b)
        assert L.zng_adler32_combine(aa, ab, lb) == pyzlib.adler32(a + b), (la, lb)
    # lengths far beyond anything that is ever materialised: the algebra must hold (x^(8n) by square and multiply)
    c1, c2 = 0x12345678, 0x9abcdef0
    n1, n2 = (1 << 40) + 12345, (1 << 33) + 7
    left = L.zng_crc32_combine(L.zng_crc32_combine(c1, c2, n1), 0xdeadbeef, n2)
    right = L.zng_crc32_combine(c1, L.zng_crc32_combine(c2, 0xdeadbeef, n2), n1 + n2)
    assert left == right                                   # associativity over 1 TiB-scale lengths
    assert L.zng_adler32_combine(1, 1, -1) == 0xffffffff   # adler32.c:37-38: negative length


def test_bounds_and_version(pkg, L):
    assert L.zlibng_version().decode().startswith("2.2.2")
    for n in (0, 1, 65535, 65536, 65537, 1 << 20, (1 << 30) + 5):
        b = L.zng_compressBound(n)
        assert b >= n + n // 8 + (n // 65536 + 1) * 8                      # 9 bits per byte + 8 bytes per piece
        assert L.zng_deflateBound(None, n) >= b
    assert pkg.deflate_bound(65536) % 16 == 0


def test_argument_checks_need_no_device(pkg, L):
    s = pkg.ZngStream()
    assert L.zng_deflateInit2(None, 1, 8, 15, 8, 0) == pkg.Z_STREAM_ERROR
    for args in ((1, 7, 15, 8, 0), (10, 8, 15, 8, 0), (1, 8, 16 + 16, 8, 0), (1, 8, 15, 10, 0), (1, 8, 15, 8, 5)):
        assert L.zng_deflateInit2(ctypes.byref(s), *args) == pkg.Z_STREAM_ERROR, args
    assert L.zng_deflate(ctypes.byref(s), 4) == pkg.Z_STREAM_ERROR        # no state
    assert L.zng_inflate(ctypes.byref(s), 4) == pkg.Z_STREAM_ERROR
    assert L.zng_deflateEnd(ctypes.byref(s)) == pkg.Z_STREAM_ERROR
    assert L.zng_deflateSetDictionary(ctypes.byref(s), None, 0) == pkg.Z_STREAM_ERROR


def test_without_a_device_nothing_is_computed_on_the_cpu(pkg, L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    s = pkg.ZngStream()
    r = L.zng_deflateInit2(ctypes.byref(s), 1, 8, 15, 8, 0)
    assert r == pkg.Z_MEM_ERROR and s.msg and b"CUDA" in s.msg               # loud, and no fallback
    buf = np.zeros(64, dtype=np.uint8)
    dl = ctypes.c_size_t(64)
    assert L.zng_compress2(buf.ctypes.data, ctypes.byref(dl), buf.ctypes.data, 10, 1) == pkg.Z_MEM_ERROR
    assert L.zng_b200_device_count() == 0


def test_checksum_without_a_device_aborts_instead_of_returning_a_plausible_value(pkg):
    """zng_crc32 has no way to report an error; a made-up value would end in a corrupt gzip trailer.  No device -> stderr + abort."""
    import subprocess, sys, torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    code = ("import ctypes, sys; L = ctypes.CDLL(sys.argv[1]); L.zng_crc32.restype = ctypes.c_uint32; "
            "L.zng_crc32.argtypes = [ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32]; print(L.zng_crc32(0, b'0123456789', 10))")
    p = subprocess.run([sys.executable, "-c", code, pkg.LIB_PATH], capture_output=True, text=True)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr and p.stdout.strip() == ""
