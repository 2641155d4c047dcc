"""The C11 host library (zlib-ng's own API names, include/zlib-ng.h) on the GPU: the call sequences of the
reference's example.c / minigzip.c for this path -- zng_deflateInit2 + zng_deflate(Z_FULL_FLUSH ... Z_FINISH),
zng_inflateInit2 + zng_inflate(Z_FINISH), zng_compress2 / zng_uncompress, zng_crc32 / zng_adler32 / combine --
compared with the oracle (and with the unmodified reference when oracle/_ref is present)."""
import ctypes
import struct
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg, ctx):
    return pkg.lib()


def expected_stream(zo, data, level, wb):
    """raw / zlib / gzip stream for the pigz-style call sequence (64 KiB pieces, Z_FULL_FLUSH, Z_FINISH last)."""
    if zo.have_ref():
        return zo.ref_deflate_stream(data, 65536, level, wb).tobytes()
    n = data.size
    nch = (n + 65535) // 65536
    body = (nch - 1) * 65536 if nch else 0
    parts = []
    if body:
        out, sizes, _, _ = zo.port_deflate_chunks(data[:body], 65536, level, 3)
        parts += [out[i, : sizes[i]].tobytes() for i in range(len(sizes))]
    if n - body:
        out, sizes, _, _ = zo.port_deflate_chunks(data[body:], 65536, level, 4)
        parts.append(out[0, : sizes[0]].tobytes())
    else:
        parts.append(b"\x03\x00")
    raw = b"".join(parts)
    if wb < 0:
        return raw
    if wb == 15:
        return bytes([0x78, 0x01 if level < 2 else (0x5e if level < 6 else 0x9c)]) + raw + struct.pack(">I", pyzlib.adler32(data.tobytes()))
    return bytes([31, 139, 8, 0, 0, 0, 0, 0, 4 if level < 2 else 0, 3]) + raw + struct.pack("<II", pyzlib.crc32(data.tobytes()), n & 0xffffffff)


def deflate_via_api(pkg, L, data, level, wb, pieces, out_room=None):
    """Feed `data` through zng_deflate; pieces = list of (nbytes, flush).  Returns (bytes, strm)."""
    s = pkg.ZngStream()
    assert L.zng_deflateInit2(ctypes.byref(s), level, 8, wb, 8, 0) == 0
    bound = int(L.zng_deflateBound(ctypes.byref(s), data.size)) + 64
    out = np.zeros(bound, dtype=np.uint8)
    got = bytearray()
    pos = 0
    for nbytes, flush in pieces:
        s.next_in = data.ctypes.data + pos
        s.avail_in = nbytes
        pos += nbytes
        while True:
            room = out_room or bound
            s.next_out = out.ctypes.data
            s.avail_out = room
            r = L.zng_deflate(ctypes.byref(s), flush)
            got += out[: room - s.avail_out].tobytes()
            assert r in (0, 1), (r, s.msg)
            if r == 1 or (s.avail_out != 0 and s.avail_in == 0):
                break
        assert s.avail_in == 0
    assert s.total_in == data.size and s.total_out == len(got)
    adler = s.adler
    assert L.zng_deflateEnd(ctypes.byref(s)) == 0
    return bytes(got), adler


@pytest.mark.parametrize("wb", [-15, 15, 31])
@pytest.mark.parametrize("n", [0, 1, 65536, 3 * 65536 + 777, (8 << 20) + 5])
def test_zng_deflate_one_call_equals_reference_piecewise(pkg, L, zo, n, wb):
    data = synth(n, seed=n % 97 + 2)
    got, adler = deflate_via_api(pkg, L, data, 1, wb, [(n, pkg.Z_FINISH)])
    assert got == expected_stream(zo, data, 1, wb)
    assert pyzlib.decompress(got, wbits=wb) == data.tobytes()
    if wb == 15:
        assert adler == pyzlib.adler32(data.tobytes())
    if wb == 31:
        assert adler == pyzlib.crc32(data.tobytes())


@pytest.mark.parametrize("level", [2, 3, 4, 5, 6, -1])
def test_zng_deflate_every_level(pkg, L, zo, level):
    """Levels 2-6 and Z_DEFAULT_COMPRESSION (= 6, deflate.c:43) through zng_deflateInit2 / zng_deflate, zlib and gzip wrappers
    (the FLEVEL bits of the zlib header follow deflate.c:873-880)."""
    n = 3 * 65536 + 777
    data = synth(n, seed=level + 40)
    for wb in (15, 31):
        got, _ = deflate_via_api(pkg, L, data, level, wb, [(n, pkg.Z_FINISH)])
        assert got == expected_stream(zo, data, 6 if level < 0 else level, wb)
        assert pyzlib.decompress(got, wbits=wb) == data.tobytes()


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6])
def test_zng_deflateSetDictionary_dependent_chunks(pkg, L, zo, level):
    """pigz's loop on the library: per chunk zng_deflateReset + zng_deflateSetDictionary(last 32 KiB of the previous chunk) +
    zng_deflate(Z_SYNC_FLUSH / Z_FINISH) -- every chunk equals what the unmodified reference emits for the same calls on a
    fresh stream, and one call over several pieces equals the same pieces joined."""
    data = synth(5 * 65536 + 4321, seed=73)
    n = data.size
    exp, esz, _, _ = (zo.ref_deflate_chunks_primed if zo.have_ref() else zo.port_deflate_chunks_primed)(data, 65536, level, 2)
    expf, efsz, _, _ = (zo.ref_deflate_chunks_primed if zo.have_ref() else zo.port_deflate_chunks_primed)(data, 65536, level, 4)
    nch = len(esz)
    s = pkg.ZngStream()
    assert L.zng_deflateInit2(ctypes.byref(s), level, 8, -15, 8, 0) == 0
    parts = []
    for i in range(nch):
        piece = np.ascontiguousarray(data[i * 65536:(i + 1) * 65536])
        assert L.zng_deflateReset(ctypes.byref(s)) == 0
        if i:
            assert L.zng_deflateSetDictionary(ctypes.byref(s), data[i * 65536 - 40000:].ctypes.data, 40000) == 0    # only the last 32768 count
        out = np.zeros(int(L.zng_deflateBound(ctypes.byref(s), piece.size)) + 64, dtype=np.uint8)
        s.next_in = piece.ctypes.data; s.avail_in = piece.size; s.next_out = out.ctypes.data; s.avail_out = out.size
        last = i == nch - 1
        assert L.zng_deflate(ctypes.byref(s), pkg.Z_FINISH if last else 2) == (1 if last else 0)
        got = out[: s.total_out].tobytes()
        assert got == (expf[i, : efsz[i]] if last else exp[i, : esz[i]]).tobytes(), i
        parts.append(got)
    assert pyzlib.decompress(b"".join(parts), wbits=-15) == data.tobytes()
    # one call for everything behind the first chunk: the same pieces, joined
    assert L.zng_deflateReset(ctypes.byref(s)) == 0
    assert L.zng_deflateSetDictionary(ctypes.byref(s), data[65536 - 32768:].ctypes.data, 32768) == 0
    rest = np.ascontiguousarray(data[65536:])
    out = np.zeros(int(L.zng_deflateBound(ctypes.byref(s), rest.size)) + 64, dtype=np.uint8)
    s.next_in = rest.ctypes.data; s.avail_in = rest.size; s.next_out = out.ctypes.data; s.avail_out = out.size
    assert L.zng_deflate(ctypes.byref(s), pkg.Z_FINISH) == 1
    assert out[: s.total_out].tobytes() == b"".join(parts[1:])
    # what the GPU path cannot reproduce is refused, not approximated
    assert L.zng_deflateSetDictionary(ctypes.byref(s), data.ctypes.data, 32768) in (0, pkg.Z_STREAM_ERROR)
    assert L.zng_deflateEnd(ctypes.byref(s)) in (0, pkg.Z_DATA_ERROR)
    for lv, wb in ((1, 31), (1, 15), (6, 31)):
        t = pkg.ZngStream()
        assert L.zng_deflateInit2(ctypes.byref(t), lv, 8, wb, 8, 0) == 0
        assert L.zng_deflateSetDictionary(ctypes.byref(t), data.ctypes.data, 32768) == pkg.Z_STREAM_ERROR
        assert L.zng_deflateEnd(ctypes.byref(t)) == 0
    t = pkg.ZngStream()
    assert L.zng_deflateInit2(ctypes.byref(t), 1, 8, -15, 8, 0) == 0
    assert L.zng_deflateSetDictionary(ctypes.byref(t), data.ctypes.data, 1000) == pkg.Z_STREAM_ERROR        # shorter than a window
    assert L.zng_deflateEnd(ctypes.byref(t)) == 0


def test_zng_deflate_piecewise_and_small_output_windows(pkg, L, zo):
    n = 5 * 65536 + 1234
    data = synth(n, seed=11)
    exp = expected_stream(zo, data, 1, 31)
    # the reference's own call sequence: one 64 KiB piece per Z_FULL_FLUSH call, Z_FINISH on the tail
    pieces = [(65536, pkg.Z_FULL_FLUSH)] * 5 + [(1234, pkg.Z_FINISH)]
    assert deflate_via_api(pkg, L, data, 1, 31, pieces)[0] == exp
    # several pieces per call, and Z_NO_FLUSH buffering in between
    pieces = [(2 * 65536, pkg.Z_FULL_FLUSH), (65536, pkg.Z_NO_FLUSH), (2 * 65536, pkg.Z_FULL_FLUSH), (1234, pkg.Z_FINISH)]
    assert deflate_via_api(pkg, L, data, 1, 31, pieces)[0] == exp
    # output handed out through a 1000-byte window (pending output like the reference's pending_buf)
    assert deflate_via_api(pkg, L, data, 1, 31, [(n, pkg.Z_FINISH)], out_room=1000)[0] == exp


def test_zng_deflate_argument_errors(pkg, L):
    s = pkg.ZngStream()
    assert L.zng_deflateInit2(None, 1, 8, 15, 8, 0) == pkg.Z_STREAM_ERROR
    for args in ((1, 7, 15, 8, 0), (10, 8, 15, 8, 0), (1, 8, 16 + 16, 8, 0), (1, 8, 15, 10, 0), (1, 8, 15, 8, 5), (7, 8, 15, 8, 0), (9, 8, 15, 8, 0), (0, 8, 15, 8, 0), (1, 8, 12, 8, 0)):
        assert L.zng_deflateInit2(ctypes.byref(s), *args) == pkg.Z_STREAM_ERROR, args
    assert L.zng_deflateInit2(ctypes.byref(s), 1, 8, 15, 8, 0) == 0
    buf = np.zeros(64, dtype=np.uint8)
    s.next_in = buf.ctypes.data; s.avail_in = 4; s.next_out = None; s.avail_out = 64
    assert L.zng_deflate(ctypes.byref(s), pkg.Z_FINISH) == pkg.Z_STREAM_ERROR           # next_out NULL (deflate.c:823)
    s.next_out = buf.ctypes.data; s.avail_out = 0
    assert L.zng_deflate(ctypes.byref(s), pkg.Z_FINISH) == pkg.Z_BUF_ERROR              # avail_out == 0 (deflate.c:831)
    assert L.zng_deflate(ctypes.byref(s), 9) == pkg.Z_STREAM_ERROR
    assert L.zng_deflateEnd(ctypes.byref(s)) in (0, pkg.Z_DATA_ERROR)
    assert L.zng_deflateEnd(ctypes.byref(s)) == pkg.Z_STREAM_ERROR                      # already ended


def inflate_once(pkg, L, stream, wb, cap):
    s = pkg.ZngStream()
    assert L.zng_inflateInit2(ctypes.byref(s), wb) == 0
    src = np.frombuffer(stream + b"\0", dtype=np.uint8).copy()
    out = np.zeros(cap + 1, dtype=np.uint8)
    s.next_in = src.ctypes.data; s.avail_in = len(stream); s.next_out = out.ctypes.data; s.avail_out = cap
    r = L.zng_inflate(ctypes.byref(s), pkg.Z_FINISH)
    res = (r, out[: s.total_out].copy(), int(s.total_in), int(s.adler), s.msg.decode() if s.msg else None, int(s.avail_in))
    assert L.zng_inflateEnd(ctypes.byref(s)) == 0
    return res


def test_zng_inflate_oneshot_matches_oracle(pkg, L, zo):
    rng = np.random.default_rng(5)
    for trial in range(24):
        size = int(rng.integers(0, 200000))
        data = synth(size + 10, seed=trial + 3)[:size].tobytes()
        wb = [-15, 15, 31][trial % 3]
        co = pyzlib.compressobj(int(rng.choice([1, 6])), pyzlib.DEFLATED, wb)
        st = co.compress(data) + co.flush()
        variants = [(st, size + 10), (st, max(size - 3, 0)), (st[: len(st) // 2], size + 10), (st + b"next member", size)]
        s2 = bytearray(st); s2[len(s2) // 2] ^= 4
        variants.append((bytes(s2), size + 10))
        for stream, cap in variants:
            g = inflate_once(pkg, L, stream, wb, cap)
            e = zo.port_inflate(stream, wb, cap)
            assert g[0] == e[0] and g[4] == e[4], (trial, wb, cap, g[0], g[4], e[0], e[4])
            if e[0] == 1:
                assert np.array_equal(g[1], e[1]) and g[2] == e[2] and g[3] == e[3]
                assert g[5] == len(stream) - e[2]                       # unused input stays with the caller


def test_zng_inflate_incremental(pkg, L):
    data = synth(300000, seed=8).tobytes()
    st = pyzlib.compress(data, 6)
    src = np.frombuffer(st, dtype=np.uint8).copy()
    s = pkg.ZngStream()
    assert L.zng_inflateInit2(ctypes.byref(s), 15) == 0
    out = np.zeros(70000, dtype=np.uint8)
    got = bytearray()
    pos, r = 0, 0
    while r != 1:
        take = min(50000, len(st) - pos)
        s.next_in = src.ctypes.data + pos; s.avail_in = take
        s.next_out = out.ctypes.data; s.avail_out = out.size
        r = L.zng_inflate(ctypes.byref(s), pkg.Z_NO_FLUSH)
        assert r in (0, 1), (r, s.msg)
        pos += take - s.avail_in
        got += out[: out.size - s.avail_out].tobytes()
    assert bytes(got) == data and s.total_out == len(data) and s.total_in == len(st) and s.adler == pyzlib.adler32(data)
    assert L.zng_inflateEnd(ctypes.byref(s)) == 0


def _inflate_pieces(pkg, L, stream, wb, piece, out_room, flush_last=False):
    """Feed `stream` to zng_inflate in `piece`-byte pieces with `out_room` bytes of output room per call (pieces / room may be
    callables of the call index).  Returns (ret, output, total_in, adler, calls, first_output_call)."""
    src = np.frombuffer(stream + b"\0", dtype=np.uint8).copy()
    s = pkg.ZngStream()
    assert L.zng_inflateInit2(ctypes.byref(s), wb) == 0
    got = bytearray()
    pos, r, calls, first_out = 0, 0, 0, None
    while r == 0 and calls < 100000:
        take = min(piece(calls) if callable(piece) else piece, len(stream) - pos)
        room = out_room(calls) if callable(out_room) else out_room
        out = np.zeros(max(room, 1), dtype=np.uint8)
        s.next_in = src.ctypes.data + pos; s.avail_in = take
        s.next_out = out.ctypes.data; s.avail_out = room
        last = flush_last and pos + take == len(stream)
        r = L.zng_inflate(ctypes.byref(s), pkg.Z_FINISH if last else pkg.Z_NO_FLUSH)
        pos += take - s.avail_in
        n_out = room - s.avail_out
        if n_out and first_out is None:
            first_out = calls
        got += out[:n_out].tobytes()
        calls += 1
        if r == pkg.Z_BUF_ERROR and take == 0 and n_out == 0:
            break                                           # no input left and nothing more to give: an incomplete stream
        if r == pkg.Z_BUF_ERROR:
            r = 0
    res = (r, bytes(got), int(s.total_in), int(s.adler), calls, first_out, s.msg.decode() if s.msg else None)
    assert L.zng_inflateEnd(ctypes.byref(s)) == 0
    return res


@pytest.mark.parametrize("wb", [-15, 15, 31])
def test_zng_inflate_piecewise_is_resumable(pkg, L, wb):
    """zng_inflate(Z_NO_FLUSH) over a member WITHOUT flush markers, fed in pieces: output arrives while the input is still being fed
    (block-boundary checkpoints of the device decoder, zng_b200_inflate_stream_feed), every byte and the check value are right, the
    bytes after the stream stay with the caller -- for piece sizes from 1 byte up and output rooms from 1 byte up."""
    rng = np.random.default_rng(wb + 40)
    data = synth(1024 * 1024 + 12345, seed=50 + abs(wb)).tobytes()
    co = pyzlib.compressobj(6, pyzlib.DEFLATED, wb)
    st = co.compress(data) + co.flush()
    chk = 0 if wb < 0 else (pyzlib.adler32(data) if wb == 15 else pyzlib.crc32(data))
    for piece, room in ((65536, 1 << 22), (10007, 4096), (lambda i: int(rng.integers(1, 9000)), lambda i: int(rng.integers(1, 70000))), (len(st), 1000)):
        r, got, tin, adler, calls, first_out, _ = _inflate_pieces(pkg, L, st + b"trailing bytes of the next member", wb, piece, room)
        assert r == 1 and got == data and tin == len(st), (wb, r, len(got), tin, len(st))
        if wb >= 0:
            assert adler == chk
        if not callable(piece) and piece < len(st) // 4:
            assert first_out is not None and first_out < calls // 2, (first_out, calls)       # piecewise, not at the end
    # small streams, every piece size
    for n in (0, 1, 5, 300, 70000):
        small = data[:n]
        co = pyzlib.compressobj(1, pyzlib.DEFLATED, wb)
        s2 = co.compress(small) + co.flush()
        for piece in ((1, 2, 7, len(s2) + 5) if n <= 300 else (97, 4096, len(s2) + 5)):
            r, got, tin, _, _, _, _ = _inflate_pieces(pkg, L, s2 + b"next", wb, piece, 100000)
            assert r == 1 and got == small and tin == len(s2), (wb, n, piece, r, tin, len(s2))


def test_zng_inflate_piecewise_errors_and_truncation(pkg, L):
    data = synth(1 << 20, seed=77).tobytes()
    st = pyzlib.compress(data, 6)
    bad = bytearray(st); bad[len(bad) // 2] ^= 0x10
    r, got, _, _, _, _, msg = _inflate_pieces(pkg, L, bytes(bad), 15, 4096, 1 << 20)
    assert r in (pkg.Z_DATA_ERROR, pkg.Z_BUF_ERROR) and (r != pkg.Z_DATA_ERROR or msg)
    # as with the reference, whole blocks decoded before the error surfaces are delivered (a flipped bit can decode "validly" for a
    # while): what is certain is that the output ahead of the damaged block is right
    k = min(len(got), 256 * 1024)
    assert k > 0 and got[:k] == data[:k]
    r, got, _, _, _, _, _ = _inflate_pieces(pkg, L, st[: len(st) // 2], 15, 4096, 1 << 20)
    assert r == pkg.Z_BUF_ERROR and data.startswith(got) and len(got) > 0     # truncated: never Z_STREAM_END
    tr = bytearray(st); tr[-1] ^= 1                                            # wrong Adler-32 in the trailer
    r, got, _, _, _, _, msg = _inflate_pieces(pkg, L, bytes(tr), 15, 65536, 1 << 21)
    assert r == pkg.Z_DATA_ERROR and msg == "incorrect data check"
    gz = pyzlib.compressobj(6, pyzlib.DEFLATED, 31); g = gz.compress(data) + gz.flush()
    tr = bytearray(g); tr[-2] ^= 1                                             # wrong ISIZE
    r, _, _, _, _, _, msg = _inflate_pieces(pkg, L, bytes(tr), 31, 65536, 1 << 21)
    assert r == pkg.Z_DATA_ERROR and msg == "incorrect length check"


def test_zng_inflate_piecewise_bounded_memory(pkg, L, monkeypatch):
    """The resumable decoder keeps 32 KiB of history plus what the caller has not taken yet, not the whole output: delivered output
    is dropped from the device buffer (threshold lowered here so that it happens many times), the check value is carried from call
    to call, and a caller who drains slowly is refused more input (like the reference with avail_out == 0) instead of piling up
    output.  Same bytes, same CRC-32 / Adler-32 / ISIZE verdict, same total_in."""
    monkeypatch.setenv("ZNG_B200_INFLATE_COMPACT", "65536")
    monkeypatch.setenv("ZNG_B200_INFLATE_BACKLOG", "262144")
    data = synth(3 * 1024 * 1024 + 777, seed=91).tobytes()
    for wb in (31, 15, -15):
        co = pyzlib.compressobj(6, pyzlib.DEFLATED, wb)
        st = co.compress(data) + co.flush()
        chk = 0 if wb < 0 else (pyzlib.adler32(data) if wb == 15 else pyzlib.crc32(data))
        for piece, room in ((65536, 1 << 22), (30000, 20000), (len(st), 50000), (50000, 3000)):
            r, got, tin, adler, calls, _, msg = _inflate_pieces(pkg, L, st + b"tail", wb, piece, room)
            assert r == 1 and got == data and tin == len(st), (wb, piece, room, r, msg, len(got), tin, len(st), calls)
            if wb >= 0:
                assert adler == chk
    co = pyzlib.compressobj(6, pyzlib.DEFLATED, 31); g = bytearray(co.compress(data) + co.flush()); g[-6] ^= 1
    r, _, _, _, _, _, msg = _inflate_pieces(pkg, L, bytes(g), 31, 65536, 1 << 20)
    assert r == pkg.Z_DATA_ERROR and msg == "incorrect data check"


def test_zng_inflate_piecewise_fuzz_against_cpython(pkg, L):
    """Deterministic fuzz of the piecewise path: valid, truncated and damaged streams (all three wrappers, levels 1 / 6 / 9, stored
    blocks included) fed in random pieces with random output room, next to CPython's zlib fed the same pieces.  A stream CPython
    finishes must finish here with the same bytes and the same consumed length; one it rejects must end in Z_DATA_ERROR (or
    Z_NEED_DICT) here; one it leaves open must stay open (Z_BUF_ERROR when nothing is left to feed) -- and whatever was delivered
    is a prefix of what CPython produced (whole blocks here, single symbols there)."""
    rng = np.random.default_rng(20261019)
    base = synth(400 * 1024, seed=123).tobytes()
    noise = rng.integers(0, 256, size=70000, dtype=np.uint8).tobytes()
    n_done = n_err = n_open = 0
    for case in range(90):
        wb = (-15, 15, 31)[case % 3]
        level = (1, 6, 9)[(case // 3) % 3]
        n = int(rng.integers(0, 200000))
        o = int(rng.integers(0, len(base) - n))
        data = base[o:o + n] if case % 7 else (base[o:o + n // 2] + noise[: n // 3])     # every 7th: stored blocks inside
        co = pyzlib.compressobj(level, pyzlib.DEFLATED, wb)
        st = bytearray(co.compress(data) + co.flush())
        kind = case % 5                                                             # 0, 1: intact; 2: truncated; 3: bit flip; 4: bytes overwritten
        if kind == 2 and len(st) > 4:
            st = st[: int(rng.integers(1, len(st)))]
        elif kind == 3 and len(st) > 4:
            k = int(rng.integers(0, len(st))); st[k] ^= 1 << int(rng.integers(0, 8))
        elif kind == 4 and len(st) > 12:
            k = int(rng.integers(2, len(st) - 8)); st[k:k + 4] = bytes(rng.integers(0, 256, size=4, dtype=np.uint8))
        tail = b"" if kind == 2 else b"bytes behind the stream"
        stream = bytes(st) + tail
        hi = int(rng.integers(2, 9000)) if len(stream) < 3000 else int(rng.integers(300, 9000))
        sizes = [int(x) for x in rng.integers(1, hi, size=4096)]
        rooms = [int(x) for x in rng.integers(1, 60000, size=4096)]
        # CPython over the same pieces
        d = pyzlib.decompressobj(wb)
        ref_out, ref_err, pos = bytearray(), False, 0
        try:
            i = 0
            while pos < len(stream) and not d.eof:
                sz = sizes[i % 4096]; i += 1
                ref_out += d.decompress(stream[pos:pos + sz]); pos += sz
        except pyzlib.error:
            ref_err = True
        r, got, tin, _, calls, _, msg = _inflate_pieces(pkg, L, stream, wb, lambda i: sizes[i % 4096], lambda i: rooms[i % 4096])
        ctxt = (case, wb, level, kind, len(data), len(stream), r, msg, len(got), tin, calls)
        if ref_err:
            assert r in (pkg.Z_DATA_ERROR, pkg.Z_NEED_DICT), ctxt
            k = min(len(got), len(ref_out))                          # (CPython loses the output of the call that raises)
            assert got[:k] == bytes(ref_out[:k]), ctxt
            n_err += 1
        elif d.eof:
            assert r == 1 and got == bytes(ref_out) and tin == min(pos, len(stream)) - len(d.unused_data), ctxt
            n_done += 1
        else:
            assert r == pkg.Z_BUF_ERROR and bytes(ref_out).startswith(got), ctxt
            n_open += 1
    assert n_done >= 30 and n_err >= 10 and n_open >= 5, (n_done, n_err, n_open)


def test_zng_inflate_incremental_is_not_quadratic(pkg, L):
    """The common loop "read 64 KiB, zng_inflate(Z_NO_FLUSH)" over a gzip member without flush markers: every call resumes at the last
    block boundary (host/zng_inflate.c -> zng_b200_inflate_stream_feed), so 48 MiB of compressed input costs one pass, not 768 full decodes."""
    import time
    raw = np.random.default_rng(3).integers(0, 256, size=48 << 20, dtype=np.uint8).tobytes()   # incompressible: stored blocks
    co = pyzlib.compressobj(1, pyzlib.DEFLATED, 31)
    st = co.compress(raw) + co.flush()
    src = np.frombuffer(st, dtype=np.uint8).copy()
    s = pkg.ZngStream()
    assert L.zng_inflateInit2(ctypes.byref(s), 31) == 0
    out = np.zeros(len(raw) + 64, dtype=np.uint8)
    s.next_out = out.ctypes.data; s.avail_out = out.size
    pos, r, calls = 0, 0, 0
    t0 = time.perf_counter()
    while r != 1:
        take = min(65536, len(st) - pos)
        s.next_in = src.ctypes.data + pos; s.avail_in = take
        r = L.zng_inflate(ctypes.byref(s), pkg.Z_NO_FLUSH)
        assert r in (0, 1), (r, s.msg)
        pos += take - s.avail_in
        calls += 1
        assert calls < 2000
    dt = time.perf_counter() - t0
    assert s.total_out == len(raw) and out[: len(raw)].tobytes() == raw and s.total_in == len(st)
    assert dt < 60, f"{calls} calls took {dt:.1f} s"
    assert L.zng_inflateEnd(ctypes.byref(s)) == 0


def test_uncompress_moves_only_the_bytes_it_produced(pkg, L):
    """A generous destLen: bytes past the decoded length stay as the caller left them (no stale device bytes travel)."""
    big = synth(3 << 20, seed=77)
    comp = pyzlib.compress(big.tobytes(), 1)
    dst = np.full(4 << 20, 0xA5, dtype=np.uint8)
    dlen = ctypes.c_size_t(dst.size)
    assert L.zng_uncompress(dst.ctypes.data, ctypes.byref(dlen), np.frombuffer(comp, dtype=np.uint8).ctypes.data, len(comp)) == 0
    small = b"tiny payload" * 3
    comp2 = np.frombuffer(pyzlib.compress(small, 6), dtype=np.uint8).copy()
    dst = np.full(4 << 20, 0xA5, dtype=np.uint8)
    dlen = ctypes.c_size_t(dst.size)
    assert L.zng_uncompress(dst.ctypes.data, ctypes.byref(dlen), comp2.ctypes.data, comp2.size) == 0
    assert dlen.value == len(small) and dst[: len(small)].tobytes() == small
    assert (dst[len(small):] == 0xA5).all()


@pytest.mark.slow
def test_adler32_of_a_large_zlib_member(pkg, L):
    """1.2 GiB of 0xFF in ONE zlib member without flush markers: the position-weighted Adler-32 sums of the member decoder pass
    2^64 unless they are reduced first (csrc/inflate.cu warp_check)."""
    n = 20 * (64 << 20)                     # 1.25 GiB
    co = pyzlib.compressobj(1, pyzlib.DEFLATED, 15)
    piece = b"\xff" * (64 << 20)
    parts, a = [], 1
    for _ in range(n // len(piece)):
        parts.append(co.compress(piece)); a = pyzlib.adler32(piece, a)
    parts.append(co.flush())
    st = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    dst = np.zeros(n, dtype=np.uint8)
    dlen = ctypes.c_size_t(n)
    assert L.zng_uncompress(dst.ctypes.data, ctypes.byref(dlen), st.ctypes.data, st.size) == 0
    assert dlen.value == n and int(dst.min()) == 0xFF


def test_compress_uncompress_roundtrip(pkg, L, zo):
    data = synth(5 * 65536 + 321, seed=21)
    cap = int(L.zng_compressBound(data.size))
    comp = np.zeros(cap, dtype=np.uint8)
    clen = ctypes.c_size_t(cap)
    assert L.zng_compress2(comp.ctypes.data, ctypes.byref(clen), data.ctypes.data, data.size, 1) == 0
    assert comp[: clen.value].tobytes() == expected_stream(zo, data, 1, 15)
    back = np.zeros(data.size, dtype=np.uint8)
    blen = ctypes.c_size_t(data.size)
    assert L.zng_uncompress(back.ctypes.data, ctypes.byref(blen), comp.ctypes.data, clen.value) == 0
    assert blen.value == data.size and np.array_equal(back, data)
    blen = ctypes.c_size_t(1000)                                        # too small: Z_BUF_ERROR (uncompr.c)
    assert L.zng_uncompress(back.ctypes.data, ctypes.byref(blen), comp.ctypes.data, clen.value) == pkg.Z_BUF_ERROR
    blen = ctypes.c_size_t(data.size)                                   # truncated input: Z_DATA_ERROR
    assert L.zng_uncompress(back.ctypes.data, ctypes.byref(blen), comp.ctypes.data, clen.value // 2) == pkg.Z_DATA_ERROR


def test_checksum_api(pkg, L, zo, golden):
    seen = 0
    for v in golden("kat_crc32.json")["vectors"]:
        if v["data_hex"] is None:                          # the reference's NULL-buffer vectors (test_crc32.cc:30-35)
            assert L.zng_crc32_z(v["init"], None, v["len"]) == v["expect"], v
            continue
        buf = np.frombuffer(bytes.fromhex(v["data_hex"]) + b"\0", dtype=np.uint8).copy()
        assert L.zng_crc32_z(v["init"], buf.ctypes.data, buf.size - 1) == v["expect"], v
        assert L.zng_crc32(v["init"], buf.ctypes.data, buf.size - 1) == v["expect"], v
        seen += 1
    assert seen >= 50
    for v in golden("kat_adler32.json")["vectors"]:
        if v["data_hex"] is None:
            if v["len"] != 1:
                assert L.zng_adler32_z(v["init"], None, v["len"]) == v["expect"], v
            continue
        buf = np.frombuffer(bytes.fromhex(v["data_hex"]) + b"\0", dtype=np.uint8).copy()
        assert L.zng_adler32_z(v["init"], buf.ctypes.data, buf.size - 1) == v["expect"], v
    assert L.zng_crc32_z(123, None, 10) == 0 and L.zng_adler32_z(123, None, 10) == 1      # crc32.c:28, adler32_c.c
    data = synth(3_000_001, seed=5)
    a, b = data[:1_234_567], data[1_234_567:]
    ca, cb = L.zng_crc32_z(0, a.ctypes.data, a.size), L.zng_crc32_z(0, b.ctypes.data, b.size)
    assert ca == pyzlib.crc32(a.tobytes()) and L.zng_crc32_combine(ca, cb, b.size) == pyzlib.crc32(data.tobytes())
    assert L.zng_crc32_combine_op(ca, cb, L.zng_crc32_combine_gen(b.size)) == pyzlib.crc32(data.tobytes())
    aa, ab = L.zng_adler32_z(1, a.ctypes.data, a.size), L.zng_adler32_z(1, b.ctypes.data, b.size)
    assert L.zng_adler32_combine(aa, ab, b.size) == pyzlib.adler32(data.tobytes())
    assert L.zng_crc32_z(ca, b.ctypes.data, b.size) == pyzlib.crc32(data.tobytes())         # running value continues
    assert L.zng_adler32_z(aa, b.ctypes.data, b.size) == pyzlib.adler32(data.tobytes())


def test_minigzip_cli_roundtrip(pkg, zo, tmp_path):
    """tools/minigzip_b200 (the test/minigzip.c use case on the GPU library): gzip -1 / -2 of a file equals the
    reference's stream for the same pieces, gunzip accepts it and python's gzip module too; -d restores the input,
    including two concatenated members."""
    import gzip
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "minigzip_b200")
    if not os.path.exists(exe):
        pytest.skip("minigzip_b200 not built")
    data = synth(7 * 65536 + 4321, seed=19)
    src = tmp_path / "in.bin"
    src.write_bytes(data.tobytes())
    for level in (1, 2, 3, 4, 5, 6):
        comp = subprocess.run([exe, f"-{level}", str(src)], stdout=subprocess.PIPE, check=True).stdout
        assert comp == expected_stream(zo, data, level, 31)
        assert gzip.decompress(comp) == data.tobytes()
        back = subprocess.run([exe, "-d"], input=comp, stdout=subprocess.PIPE, check=True).stdout
        assert back == data.tobytes()
        two = subprocess.run([exe, "-d"], input=comp + comp, stdout=subprocess.PIPE, check=True).stdout
        assert two == data.tobytes() * 2
    other = gzip.compress(data.tobytes(), 6)                 # a member some other encoder wrote
    assert subprocess.run([exe, "-d"], input=other, stdout=subprocess.PIPE, check=True).stdout == data.tobytes()
    bad = bytearray(other); bad[len(bad) // 2] ^= 1
    r = subprocess.run([exe, "-d"], input=bytes(bad), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"zng_inflate: -3" in r.stderr
