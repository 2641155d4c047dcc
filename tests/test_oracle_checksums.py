"""The oracle's checksum restatement against the reference's own known-answer tests
(test/test_crc32.cc:29-183, test/test_adler32.cc:202-345, extracted into tests/golden/ by
make_golden.py) and, when oracle/_ref is present, against the unmodified reference live."""
import numpy as np
import pytest


def _data(v):
    return None if v["data_hex"] is None else bytes.fromhex(v["data_hex"])


def test_crc32_kats(zo, golden):
    vecs = golden("kat_crc32.json")["vectors"]
    assert len(vecs) >= 100
    for v in vecs:
        d = _data(v)
        # the reference harness (test_crc32.cc:186-199): NULL -> 0, len 0 -> the seed itself
        if d is None:
            got = int(zo.port().zo_crc32(v["init"], None, v["len"]))
        elif v["len"] == 0:
            one = np.zeros(1, dtype=np.uint8)            # a valid pointer with length 0
            got = int(zo.port().zo_crc32(v["init"], one.ctypes.data, 0))
        else:
            got = zo.port_crc32(d, v["init"])
        assert got == v["expect"], v


def test_adler32_kats(zo, golden):
    vecs = golden("kat_adler32.json")["vectors"]
    assert len(vecs) >= 100
    for v in vecs:
        d = _data(v)
        if d is None:
            got = int(zo.port().zo_adler32(v["init"], None, v["len"]))
        else:
            got = zo.port_adler32(d, v["init"])
        assert got == v["expect"], v


def test_compare256_property(zo):
    # test/test_compare256.cc:25-51: a mismatch planted at every index 0..255, and none
    a = np.full(512, ord("a"), dtype=np.uint8)
    for i in range(257):
        b = a.copy()
        if i < 256:
            b[i] = ord("b")
        assert int(zo.port().zo_compare256(a.ctypes.data, b.ctypes.data)) == i


def test_checksums_vs_reference_live(zo, pkg):
    if not zo.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for n in (0, 1, 15, 16, 17, 63, 64, 65, 5551, 5552, 5553, 65536, 1 << 20):
        d = rng.integers(0, 256, size=n, dtype=np.uint8)
        for init in (0, 1, 0xdeadbeef):
            assert zo.port_crc32(d, init) == zo.ref_crc32(d, init)
            assert zo.port_adler32(d, init) == zo.ref_adler32(d, init)
    R, P = zo.ref(), zo.port()
    for c1, c2, ln in ((0, 0, 0), (0x12345678, 0x9abcdef0, 1), (0xffffffff, 1, 65536), (7, 9, (1 << 36) + 12345), (0xcafe, 0, 5)):
        assert int(R.zng_crc32_combine(c1, c2, ln)) == int(P.zo_crc32_combine(c1, c2, ln))
        assert int(R.zng_crc32_combine_gen(ln)) == int(P.zo_crc32_combine_gen(ln))
        assert int(R.zng_crc32_combine_op(c1, c2, int(R.zng_crc32_combine_gen(ln)))) == int(R.zng_crc32_combine(c1, c2, ln))
        assert int(R.zng_adler32_combine(c1, c2, ln)) == int(P.zo_adler32_combine(c1, c2, ln))
    assert int(P.zo_adler32_combine(1, 1, -1)) == 0xffffffff


def test_combine_matches_concatenation(zo):
    rng = np.random.default_rng(6)
    a = rng.integers(0, 256, size=70001, dtype=np.uint8)
    b = rng.integers(0, 256, size=12345, dtype=np.uint8)
    ab = np.concatenate([a, b])
    P = zo.port()
    assert int(P.zo_crc32_combine(zo.port_crc32(a), zo.port_crc32(b), b.size)) == zo.port_crc32(ab)
    assert int(P.zo_adler32_combine(zo.port_adler32(a), zo.port_adler32(b), b.size)) == zo.port_adler32(ab)
