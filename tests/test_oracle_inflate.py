"""The oracle's inflate restatement (oracle/zo_inflate.c = one zng_inflate(Z_FINISH) call on a whole
stream) against the reference's hand-written bitstreams (test/infcover.c inf() calls, extracted into
tests/golden/inflate_kat.json) and, when oracle/_ref is present, against the unmodified reference on
valid, truncated and bit-flipped streams (return code, strm->msg, output, bytes consumed, check value)."""
import struct
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest


def kat_streams(golden):
    for v in golden("inflate_kat.json")["vectors"]:
        yield v, bytes(int(x, 16) for x in v["hex"].split())


def test_infcover_vectors(zo, golden):
    """infcover expects the code of the FIRST zng_inflate(Z_NO_FLUSH) call; for whole-input vectors
    (step 0) that is Z_OK exactly when the stream needs more input or output, which one Z_FINISH call
    reports as Z_BUF_ERROR (inflate.c:1197-1199)."""
    seen = 0
    for v, st in kat_streams(golden):
        if v["step"] != 0 or not isinstance(v["err"], int):
            continue
        if v["what"] == "bad window size":
            assert zo.port_inflate(st, v["win"], 16)[0] == -2          # zng_inflateInit2 rejects it
            continue
        r = zo.port_inflate(st, v["win"], v["len"])
        want = v["err"]
        if want == 0:
            want = -5
        assert r[0] == want, (v, r[0], r[4])
        seen += 1
    assert seen >= 15


def gzip_member(payload: bytes, level=6, extra=None, name=None, comment=None, hcrc=False):
    flg = (4 if extra is not None else 0) | (8 if name is not None else 0) | (16 if comment is not None else 0) | (2 if hcrc else 0)
    h = bytes([0x1f, 0x8b, 8, flg, 1, 2, 3, 4, 0, 3])
    if extra is not None:
        h += struct.pack("<H", len(extra)) + extra
    if name is not None:
        h += name + b"\0"
    if comment is not None:
        h += comment + b"\0"
    if hcrc:
        h += struct.pack("<H", pyzlib.crc32(h) & 0xffff)
    co = pyzlib.compressobj(level, pyzlib.DEFLATED, -15)
    body = co.compress(payload) + co.flush()
    return h + body + struct.pack("<II", pyzlib.crc32(payload), len(payload) & 0xffffffff)


def test_gzip_header_fields(zo):
    payload = b"hello hello hello hello, header fields " * 20
    for kw in (dict(), dict(extra=b"ab\x03\x00xyz"), dict(name=b"file.txt"), dict(comment=b"a comment"),
               dict(hcrc=True), dict(extra=b"", name=b"n", comment=b"c", hcrc=True)):
        st = gzip_member(payload, **kw)
        r = zo.port_inflate(st, 31, len(payload))
        assert r[0] == 1 and r[1].tobytes() == payload and r[2] == len(st) and r[3] == pyzlib.crc32(payload), kw
        assert zo.port_inflate(st, 47, len(payload))[0] == 1
    st = bytearray(gzip_member(payload, name=b"x", hcrc=True))
    st[13] ^= 1                                           # break the header crc
    r = zo.port_inflate(bytes(st), 31, len(payload))
    assert r[0] == -3 and r[4] == "header crc mismatch"
    st = bytearray(gzip_member(payload))
    st[-5] ^= 0x10
    assert zo.port_inflate(bytes(st), 31, len(payload))[4] == "incorrect data check"
    st = bytearray(gzip_member(payload))
    st[-1] ^= 0x10
    assert zo.port_inflate(bytes(st), 31, len(payload))[4] == "incorrect length check"


def test_port_vs_reference_live(pkg, zo, golden):
    if not zo.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    n = [0]

    def cmp(stream, wb, cap, tag):
        a = zo.port_inflate(stream, wb, cap)
        b = zo.ref_inflate(stream, wb, cap)
        n[0] += 1
        assert a[0] == b[0] and a[4] == b[4], (tag, wb, cap, a[0], a[4], b[0], b[4])
        if a[0] == 1:
            assert np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3], tag

    for v, st in kat_streams(golden):
        for cap in (v["len"], 1000):
            cmp(st, v["win"], cap, v["what"])
    payload = b"hello hello hello hello, header fields " * 20
    for kw in (dict(extra=b"ab\x03\x00xyz"), dict(name=b"file.txt", comment=b"c"), dict(hcrc=True, name=b"q")):
        st = gzip_member(payload, **kw)
        cmp(st, 31, len(payload), "gzip hdr")
        cmp(st[:14], 31, len(payload), "gzip hdr trunc")
        for k in range(10, min(len(st), 40)):
            s2 = bytearray(st); s2[k] ^= 0x41
            cmp(bytes(s2), 47, len(payload), "gzip hdr flip")
    for trial in range(120):
        size = int(rng.integers(0, 20000))
        data = synth(65536 * 2, seed=trial + 1)[int(rng.integers(0, 60000)):][:size].tobytes()
        lvl = int(rng.choice([0, 1, 6, 9]))
        wb = [-15, 15, 31, 47][trial % 4]
        cwb = 31 if wb == 47 and trial % 8 < 4 else (15 if wb == 47 else wb)
        co = pyzlib.compressobj(lvl, pyzlib.DEFLATED, cwb, 8, int(rng.choice([0, 1, 2, 3, 4])))
        st = co.compress(data) + co.flush()
        cmp(st, wb, size + 10, "valid")
        cmp(st, wb, max(size - 5, 0), "short-out")
        cmp(st[: len(st) // 2], wb, size + 10, "trunc")
        for k in range(6):
            if not st:
                break
            s2 = bytearray(st)
            s2[int(rng.integers(0, len(s2)))] ^= 1 << int(rng.integers(0, 8))
            cmp(bytes(s2), wb, size + 10, "flip")
    # what the reference's own deflate produces for the pigz-style call sequence, levels 1 and 2
    for level in (1, 2):
        data = synth(3 * 65536 + 777, seed=level)
        for wb in (-15, 15, 31):
            st = zo.ref_deflate_stream(data, 65536, level, wb)
            cmp(st, wb, data.size, "ref stream")
    assert n[0] > 1000
