"""The host-buffer entry point (what zng_deflate of the host library calls): pipelined H2D ->
K1 -> gather -> D2H, compared with the oracle's concatenated chunk stream."""
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

pytestmark = pytest.mark.gpu


def oracle_stream(zo, data, chunk, level, final):
    n = data.size
    nch = (n + chunk - 1) // chunk
    parts = []
    if final:
        body = (nch - 1) * chunk if nch else 0
        if body:
            out, sizes, _, _ = zo.port_deflate_chunks(data[:body], chunk, level, 3)
            parts += [out[i, : sizes[i]].tobytes() for i in range(len(sizes))]
        if n - body:
            out, sizes, _, _ = zo.port_deflate_chunks(data[body:], chunk, level, 4)
            parts.append(out[0, : sizes[0]].tobytes())
        else:
            parts.append(b"\x03\x00")
    else:
        out, sizes, _, _ = zo.port_deflate_chunks(data, chunk, level, 3)
        parts += [out[i, : sizes[i]].tobytes() for i in range(len(sizes))]
    return b"".join(parts)


@pytest.mark.parametrize("n,final", [(0, True), (0, False), (1, True), (65536, True), (65536, False), (5 * 65536 + 99, True),
                                     (5 * 65536 + 99, False), ((40 << 20) + 12345, True), ((100 << 20), False)])
def test_deflate_host_matches_oracle_stream(pkg, ctx, zo, n, final):
    data = synth(n, seed=n % 1000 + 1)
    cap = pkg.deflate_bound(65536) * ((n + 65535) // 65536 + 1)
    out = np.empty(cap, dtype=np.uint8)
    out_len, crc, adler = ctx.deflate_host(data, n, 65536, 1, final, out, cap)
    exp = oracle_stream(zo, data, 65536, 1, final)
    assert out_len == len(exp)
    assert out[:out_len].tobytes() == exp
    assert crc == pyzlib.crc32(data.tobytes()) and adler == pyzlib.adler32(data.tobytes())
    if final:
        assert pyzlib.decompress(out[:out_len].tobytes(), wbits=-15) == data.tobytes()


def test_deflate_host_output_too_small(pkg, ctx):
    data = synth(4 * 65536, seed=3)
    out = np.empty(1000, dtype=np.uint8)
    with pytest.raises(pkg.ZngB200Error) as ei:
        ctx.deflate_host(data, data.size, 65536, 1, True, out, 1000)
    assert ei.value.code == pkg.Z_BUF_ERROR


def test_more_chunks_than_one_batch(pkg, ctx, zo):
    """More than 16384 chunks in one call: the chunk compressor runs several launches that share the token scratch
    (small chunks keep the test cheap); level 2's later batches still see the stream bytes in front of them."""
    import torch
    for level, chunk, nch in ((1, 2048, 40000), (2, 2048, 20000), (2, 65536, 16384 + 3)):
        n = chunk * nch - 777
        data = synth(n, seed=level * 7 + 1)
        d_in = torch.from_numpy(data).to(f"cuda:{ctx.device}")
        slots, stride, sizes, crcs, adlers = ctx.alloc_chunk_outputs(n, chunk)
        ctx.deflate_chunks(d_in, n, chunk, level, 3, slots, stride, sizes, crcs, adlers)
        torch.cuda.synchronize()
        hs = sizes.cpu().numpy().view(np.uint32)[:nch]
        hslots = slots.cpu().numpy().reshape(-1, stride)
        pick = np.concatenate([np.arange(0, 4), np.arange(16380, 16390), np.arange(nch - 3, nch), np.random.default_rng(1).choice(nch, 40, replace=False)])
        for ci in np.unique(pick[pick < nch]):
            lo = ci * chunk
            piece = data[lo: min(lo + chunk, n)]
            if level == 2 and chunk == 65536 and piece.size < chunk:
                # a short last chunk sees the previous chunk's window bytes: compress it in stream context with the oracle
                exp, es, _, _ = zo.port_deflate_chunks(data[lo - chunk:], chunk, level, 3, stride, nthreads=1)
                e_bytes, e_size = exp[1], es[1]
            else:
                exp, es, _, _ = zo.port_deflate_chunks(piece, chunk, level, 3, stride, nthreads=1)
                e_bytes, e_size = exp[0], es[0]
            assert hs[ci] == e_size and np.array_equal(hslots[ci, : e_size], e_bytes[: e_size]), (level, chunk, int(ci))
        assert int(crcs.cpu().numpy().view(np.uint32)[nch - 1]) == pyzlib.crc32(data[(nch - 1) * chunk:].tobytes())


def test_two_host_threads_two_contexts(pkg, zo):
    """One context per host thread (the reference's "one zng_stream per thread"): two threads compress and inflate
    different buffers at the same time through zlib-ng's API names; results must not interfere
    (test_deflate_concurrency.cc is the reference's analogue)."""
    import ctypes
    import threading
    L = pkg.lib()
    results = {}

    def work(tid):
        data = synth((6 + tid) * 65536 + 1000 * tid, seed=40 + tid)
        for rep in range(3):
            s = pkg.ZngStream()
            assert L.zng_deflateInit2(ctypes.byref(s), 1 + (tid & 1), 8, 31, 8, 0) == 0
            comp = np.zeros(int(L.zng_deflateBound(ctypes.byref(s), data.size)) + 64, dtype=np.uint8)
            s.next_in = data.ctypes.data; s.avail_in = data.size; s.next_out = comp.ctypes.data; s.avail_out = comp.size
            assert L.zng_deflate(ctypes.byref(s), 4) == 1
            clen = int(s.total_out)
            L.zng_deflateEnd(ctypes.byref(s))
            d = pkg.ZngStream()
            assert L.zng_inflateInit2(ctypes.byref(d), 31) == 0
            back = np.zeros(data.size, dtype=np.uint8)
            d.next_in = comp.ctypes.data; d.avail_in = clen; d.next_out = back.ctypes.data; d.avail_out = back.size
            assert L.zng_inflate(ctypes.byref(d), 4) == 1
            L.zng_inflateEnd(ctypes.byref(d))
            assert np.array_equal(back, data)
            results[(tid, rep)] = pyzlib.crc32(comp[:clen].tobytes())
        results[("crc", tid)] = L.zng_crc32_z(0, data.ctypes.data, data.size) == pyzlib.crc32(data.tobytes())

    ts = [threading.Thread(target=work, args=(t,)) for t in range(3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for tid in range(3):
        assert results[("crc", tid)]
        assert results[(tid, 0)] == results[(tid, 1)] == results[(tid, 2)]


def test_streamed_and_slab_pipelines_agree(pkg, monkeypatch):
    """The streamed path (one persistent parse kernel fed by the copy engine; default for level 1, opt-in for levels 2-6
    with ZNG_B200_STREAMED=2) and the slab pipeline (ZNG_B200_STREAMED=0) must write the same bytes, checksums included:
    ragged tail, Z_FINISH and flush-only endings, more than one output slab."""
    n = (72 << 20) + 4321
    data = synth(n, seed=55)
    cap = n + n // 8 + (n // 65536 + 1) * 8 + 64
    got = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("ZNG_B200_STREAMED", mode)
        c = pkg.Context(0)                                   # the knob is read when the context is created
        for level in (1, 2, 6):
            for final in (True, False):
                out = np.zeros(cap, dtype=np.uint8)
                ol, crc, ad = c.deflate_host(data, n, 65536, level, final, out, cap)
                got[(mode, level, final)] = (ol, crc, ad, out[:ol].copy())
        c.close()
    for level in (1, 2, 6):
        for final in (True, False):
            a, b = got[("0", level, final)], got[("2", level, final)]
            assert a[:3] == b[:3] and np.array_equal(a[3], b[3]), (level, final)
    stream = got[("2", 1, True)][3].tobytes()
    assert pyzlib.decompress(stream, wbits=-15) == data.tobytes()
    assert got[("2", 1, True)][1] == pyzlib.crc32(data.tobytes()) and got[("2", 1, True)][2] == pyzlib.adler32(data.tobytes())


def test_streamed_path_output_too_small_then_recovers(pkg, ctx, zo):
    """Z_BUF_ERROR from the streamed path (the output runs out in the middle of the drain) must leave the context usable."""
    n = (40 << 20) + 777
    data = synth(n, seed=8)
    small = np.zeros(n // 4, dtype=np.uint8)
    with pytest.raises(pkg.ZngB200Error) as ei:
        ctx.deflate_host(data, n, 65536, 1, True, small, small.size)
    assert ei.value.code == pkg.Z_BUF_ERROR
    cap = pkg.deflate_bound(65536) * ((n + 65535) // 65536 + 1)
    out = np.zeros(cap, dtype=np.uint8)
    ol, crc, ad = ctx.deflate_host(data, n, 65536, 1, True, out, cap)
    assert pyzlib.decompress(out[:ol].tobytes(), wbits=-15) == data.tobytes() and crc == pyzlib.crc32(data.tobytes())
