"""K4 (batched inflate, one warp per member) through the C-ABI against the oracle's restatement of one
zng_inflate(Z_FINISH) call: the reference's infcover bitstreams, valid / truncated / bit-flipped / short-output
streams made with several encoders, and the BASELINE config-4 shape (independent 4 KiB gzip members whose
bodies are level-1 Z_FINISH streams)."""
import struct
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

pytestmark = pytest.mark.gpu


def run_members(ctx, streams, wb, caps):
    """Inflate a list of byte strings as one batch; returns per-member (ret, out, in_used, check, msg)."""
    n = len(streams)
    in_off = np.zeros(n + 1, dtype=np.uint64)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    for i, s in enumerate(streams):
        in_off[i + 1] = in_off[i] + len(s)
        out_off[i + 1] = out_off[i] + caps[i]
    h_in = np.frombuffer(b"".join(streams) + b"\0", dtype=np.uint8).copy()
    h_out = np.zeros(int(out_off[-1]) + 1, dtype=np.uint8)
    sizes, checks, status, used, detail = ctx.inflate_members_host(h_in, in_off, wb, h_out, out_off)
    res = []
    import __graft_entry__ as ge
    pkg = ge.load_package()
    for i in range(n):
        o = int(out_off[i])
        res.append((int(status[i]), h_out[o: o + int(sizes[i])], int(used[i]), int(checks[i]), pkg.inflate_msg(int(detail[i]))))
    return res


def check_against_oracle(zo, got, streams, wb, caps, tags):
    for i, (g, st) in enumerate(zip(got, streams)):
        e = zo.port_inflate(st, wb, caps[i])
        assert g[0] == e[0] and g[4] == e[4], (tags[i], wb, caps[i], g[0], g[4], e[0], e[4])
        if e[0] == 1:
            assert np.array_equal(g[1], e[1]), tags[i]
            assert g[2] == e[2] and g[3] == e[3], (tags[i], g[2], e[2], hex(g[3]), hex(e[3]))


def test_infcover_vectors(ctx, zo, golden):
    by_win = {}
    for v in golden("inflate_kat.json")["vectors"]:
        if v["step"] != 0 or not isinstance(v["err"], int) or v["what"] == "bad window size":
            continue
        st = bytes(int(x, 16) for x in v["hex"].split())
        for cap in (v["len"], 1000):
            by_win.setdefault(v["win"], []).append((st, cap, v["what"]))
    assert sum(len(x) for x in by_win.values()) >= 30
    for wb, items in by_win.items():
        streams = [x[0] for x in items]; caps = [x[1] for x in items]; tags = [x[2] for x in items]
        got = run_members(ctx, streams, wb, caps)
        check_against_oracle(zo, got, streams, wb, caps, tags)


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


def test_gzip_header_fields(ctx, zo):
    payload = b"hello hello hello hello, header fields " * 20
    streams, tags = [], []
    for kw in (dict(), dict(extra=b"ab\x03\x00xyz"), dict(name=b"file.txt"), dict(comment=b"a comment"), dict(hcrc=True),
               dict(extra=b"", name=b"n", comment=b"c", hcrc=True)):
        st = gzip_member(payload, **kw)
        streams.append(st); tags.append(str(kw))
        streams.append(st[:14]); tags.append("trunc " + str(kw))
        for k in range(2, min(len(st), 40)):
            s2 = bytearray(st); s2[k] ^= 0x41
            streams.append(bytes(s2)); tags.append(f"flip {k} {kw}")
    for tail in (5, 1):
        st = bytearray(gzip_member(payload)); st[-tail] ^= 0x10
        streams.append(bytes(st)); tags.append(f"trailer flip {tail}")
    caps = [len(payload)] * len(streams)
    for wb in (31, 47):
        got = run_members(ctx, streams, wb, caps)
        check_against_oracle(zo, got, streams, wb, caps, tags)
    ok = run_members(ctx, [gzip_member(payload)], 31, [len(payload)])[0]
    assert ok[0] == 1 and ok[1].tobytes() == payload and ok[3] == pyzlib.crc32(payload)


@pytest.mark.parametrize("wb", [-15, 15, 31, 47])
def test_random_streams_vs_oracle(pkg, ctx, zo, wb):
    rng = np.random.default_rng(100 + wb)
    streams, caps, tags = [], [], []
    for trial in range(60):
        size = int(rng.integers(0, 30000))
        data = synth(65536 * 2, seed=trial + 1)[int(rng.integers(0, 60000)):][:size].tobytes()
        lvl = int(rng.choice([0, 1, 6, 9]))
        cwb = 31 if wb == 47 and trial % 2 else (15 if wb == 47 else wb)
        co = pyzlib.compressobj(lvl, pyzlib.DEFLATED, cwb, 8, int(rng.choice([0, 1, 2, 3, 4])))
        st = co.compress(data) + co.flush()
        streams.append(st); caps.append(size + 10); tags.append(f"valid {trial}")
        streams.append(st); caps.append(size); tags.append(f"exact {trial}")
        streams.append(st); caps.append(max(size - 5, 0)); tags.append(f"short-out {trial}")
        streams.append(st[: len(st) // 2]); caps.append(size + 10); tags.append(f"trunc {trial}")
        streams.append(st + b"trailing garbage"); caps.append(size + 10); tags.append(f"trailing {trial}")
        for k in range(6):
            s2 = bytearray(st)
            s2[int(rng.integers(0, len(s2)))] ^= 1 << int(rng.integers(0, 8))
            streams.append(bytes(s2)); caps.append(size + 10); tags.append(f"flip {trial}.{k}")
    got = run_members(ctx, streams, wb, caps)
    check_against_oracle(zo, got, streams, wb, caps, tags)
    assert sum(1 for g in got if g[0] == 1) >= 100


def test_large_members_and_overlaps(pkg, ctx, zo):
    """Members far larger than the shared-memory staging limit (direct global path), long overlapping runs."""
    datas = [synth(300000, seed=9).tobytes(), b"a" * 200000, bytes(range(256)) * 700, b"ab" * 70000 + b"xyz" * 30000,
             np.random.default_rng(1).integers(0, 256, 150000, dtype=np.uint8).tobytes()]
    for lvl in (1, 6):
        streams = [pyzlib.compress(d, lvl) for d in datas]
        caps = [len(d) for d in datas]
        got = run_members(ctx, streams, 15, caps)
        for g, d in zip(got, datas):
            assert g[0] == 1 and g[1].tobytes() == d and g[3] == pyzlib.adler32(d)
        check_against_oracle(zo, got, streams, 15, caps, [f"big{i}" for i in range(len(datas))])


def make_l1_members(zo, data, member=4096):
    """config 4: each `member`-byte slice as a gzip member whose body is what zng_deflate(level 1, Z_FINISH) emits
    (= minigzip -1 of the slice: 10-byte header with XFL 4 / OS 3, raw deflate, CRC32, ISIZE)."""
    out, sizes, crcs, _ = zo.port_deflate_chunks(data, member, 1, 4)
    hdr = bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4, 3])
    parts = []
    for i in range(len(sizes)):
        ln = min(member, data.size - i * member)
        parts.append(hdr + out[i, : sizes[i]].tobytes() + struct.pack("<II", int(crcs[i]), ln))
    return parts


def test_config4_members_device_resident(pkg, ctx, zo):
    import torch
    n_members = 4096
    data = synth(n_members * 4096, seed=4)
    parts = make_l1_members(zo, data)
    assert pyzlib.decompress(parts[7], wbits=31) == data[7 * 4096: 8 * 4096].tobytes()
    in_off = np.zeros(n_members + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([len(p) for p in parts])
    out_off = np.arange(n_members + 1, dtype=np.int64) * 4096
    dev = torch.device("cuda", ctx.device)
    d_in = torch.from_numpy(np.frombuffer(b"".join(parts) + b"\0" * 16, dtype=np.uint8).copy()).to(dev)
    d_io = torch.from_numpy(in_off).to(dev); d_oo = torch.from_numpy(out_off).to(dev)
    d_out = torch.zeros(n_members * 4096, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(n_members, dtype=torch.int32, device=dev); checks = torch.zeros_like(sizes)
    status = torch.zeros_like(sizes); used = torch.zeros_like(sizes); detail = torch.zeros_like(sizes)
    ctx.inflate_members(d_in, d_io, n_members, 31, d_out, d_oo, sizes, checks, status, used, detail)
    torch.cuda.synchronize()
    assert bool((status == 1).all()) and bool((sizes == 4096).all()) and bool((detail == 0).all())
    assert np.array_equal(d_out.cpu().numpy(), data)
    crc_exp = np.array([pyzlib.crc32(data[i * 4096:(i + 1) * 4096].tobytes()) for i in range(n_members)], dtype=np.uint32)
    assert np.array_equal(checks.cpu().numpy().view(np.uint32), crc_exp)
    assert np.array_equal(used.cpu().numpy().astype(np.int64), np.diff(in_off))


def test_argument_validation(pkg, ctx):
    one = np.zeros(8, dtype=np.uint8)
    off = np.array([0, 4], dtype=np.uint64)
    for wb in (-16, 7, 64):
        with pytest.raises(pkg.ZngB200Error) as ei:
            ctx.inflate_members_host(one, off, wb, one, off) if False else ctx._check(
                pkg.lib().zng_b200_inflate_members(ctx._h, 1, 1, 1, wb, 1, 1, 1, 0, 1, 0, 0, 0))
        assert ei.value.code == pkg.Z_STREAM_ERROR


# ---- parallel inflate of one flush-delimited stream ------------------------------------------
def _stream_cases(pkg, zo):
    import zlib as z
    data = synth(40 * 65536 + 12345, seed=55)
    raw = data.tobytes()
    cases = []
    for level in (1, 2):
        for wb in (-15, 15, 31):
            st = zo.ref_deflate_stream(data, 65536, level, wb).tobytes() if zo.have_ref() else None
            if st is None:                                  # no reference build on this box: python zlib with full flushes
                co = z.compressobj(level, z.DEFLATED, wb)
                st = b"".join(co.compress(raw[i:i + 65536]) + co.flush(z.Z_FULL_FLUSH) for i in range(0, len(raw), 65536)) + co.flush()
            cases.append((f"ref L{level} wb{wb}", st, wb, raw))
    # pigz-like: 128 KiB blocks, level 6, sync-flush markers only at some block ends, and a stream without any marker
    co = z.compressobj(6, z.DEFLATED, 31)
    st = b"".join(co.compress(raw[i:i + 131072]) + co.flush(z.Z_FULL_FLUSH) for i in range(0, len(raw), 131072)) + co.flush()
    cases.append(("zlib L6 128K full", st, 31, raw))
    cases.append(("zlib L6 no markers", z.compress(raw, 6), 15, raw))
    co = z.compressobj(6, z.DEFLATED, -15)                   # Z_SYNC_FLUSH keeps the window: segments are NOT independent
    st = b"".join(co.compress(raw[i:i + 65536]) + co.flush(z.Z_SYNC_FLUSH) for i in range(0, len(raw), 65536)) + co.flush()
    cases.append(("zlib L6 sync flush (dependent segments)", st, -15, raw))
    # a literal 00 00 FF FF inside the data of a stored block: a false marker in the middle of a block
    noisy = bytearray(np.random.default_rng(5).integers(0, 256, size=3 * 65536, dtype=np.uint8).tobytes())
    for k in (1000, 70000, 140000):
        noisy[k:k + 4] = b"\x00\x00\xff\xff"
    co = z.compressobj(1, z.DEFLATED, 31)
    nb = bytes(noisy) + raw[: 20 * 65536]
    st = b"".join(co.compress(nb[i:i + 65536]) + co.flush(z.Z_FULL_FLUSH) for i in range(0, len(nb), 65536)) + co.flush()
    cases.append(("false markers inside stored blocks", st, 31, nb))
    return cases


def test_stream_inflate_parallel_segments(pkg, ctx, zo):
    for tag, st, wb, raw in _stream_cases(pkg, zo):
        src = np.frombuffer(st + b"trailing bytes", dtype=np.uint8).copy()
        out = np.zeros(len(raw) + 16, dtype=np.uint8)
        status, out_len, in_used, check, detail = ctx.inflate_stream_host(src, src.size, wb, out, out.size)
        e = zo.port_inflate(st + b"trailing bytes", wb, out.size)
        assert status == 1 == e[0], (tag, status, detail)
        assert out_len == len(raw) and out[:out_len].tobytes() == raw, tag
        assert in_used == len(st) == e[2] and check == e[3], (tag, in_used, len(st), hex(check), hex(e[3]))
        # too little room: Z_BUF_ERROR, and for the parallel path the size needed
        status, need, _, _, detail = ctx.inflate_stream_host(src, src.size, wb, out, len(raw) - 1)
        assert status == pkg.Z_BUF_ERROR and (detail & 0x100), tag
        if detail & 0x400:
            assert need == len(raw), tag
        # corrupt one byte in the middle / cut the stream: same code and message as the oracle
        bad = bytearray(st); bad[len(bad) // 2] ^= 0x5a
        for variant in (bytes(bad), st[: len(st) - 3], st[: len(st) // 3]):
            v = np.frombuffer(variant + b"\0", dtype=np.uint8).copy()
            status, out_len, in_used, check, detail = ctx.inflate_stream_host(v, len(variant), wb, out, out.size)
            e = zo.port_inflate(variant, wb, out.size)
            assert status == e[0] and pkg.inflate_msg(detail) == e[4], (tag, status, e[0], pkg.inflate_msg(detail), e[4])


def test_zng_inflate_large_pigz_stream_roundtrip(pkg, ctx, zo):
    """zng_deflate -> zng_inflate through the host library on 96 MiB: the stream this library writes is decoded in
    parallel by its flush markers."""
    import ctypes
    import zlib as z
    L = pkg.lib()
    n = 96 << 20
    data = synth(n, seed=77)
    s = pkg.ZngStream()
    assert L.zng_deflateInit2(ctypes.byref(s), 1, 8, 31, 8, 0) == 0
    comp = np.zeros(int(L.zng_deflateBound(ctypes.byref(s), n)) + 64, dtype=np.uint8)
    s.next_in = data.ctypes.data; s.avail_in = n; s.next_out = comp.ctypes.data; s.avail_out = comp.size
    assert L.zng_deflate(ctypes.byref(s), pkg.Z_FINISH) == 1
    clen = int(s.total_out)
    assert L.zng_deflateEnd(ctypes.byref(s)) == 0
    d = pkg.ZngStream()
    assert L.zng_inflateInit2(ctypes.byref(d), 31) == 0
    back = np.zeros(n, dtype=np.uint8)
    d.next_in = comp.ctypes.data; d.avail_in = clen; d.next_out = back.ctypes.data; d.avail_out = n
    assert L.zng_inflate(ctypes.byref(d), pkg.Z_FINISH) == 1
    assert d.total_out == n and d.total_in == clen and d.avail_in == 0 and d.adler == z.crc32(data.tobytes())
    assert np.array_equal(back, data)
    assert L.zng_inflateEnd(ctypes.byref(d)) == 0
