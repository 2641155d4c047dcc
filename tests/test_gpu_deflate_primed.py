"""Primed-chunk parity (pigz's dependent-chunk mode) at levels 1-6: chunk i > 0 is compressed by a fresh stream primed
with zng_deflateSetDictionary(the 32768 stream bytes in front of it).  The GPU output (level 1: K1p, speculative parser on absolute
positions; levels 2-6: K2w, the reference's own window state with real slides, csrc/deflate_window.cu) must equal the oracle's
restatement (oracle/zo_deflate.c: quick_parse_primed / the window engine, pinned against the unmodified reference in
tests/test_oracle_primed.py and here, live, when oracle/_ref travelled) byte for byte, and the concatenation must inflate as ONE
raw-deflate stream."""
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

pytestmark = pytest.mark.gpu


LEVELS = [1, 2, 3, 4, 5, 6]


def gpu_primed(pkg, ctx, data, flush, level=1):
    import torch
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = data.size
    d_in = torch.from_numpy(data).to(f"cuda:{ctx.device}") if n else torch.zeros(16, dtype=torch.uint8, device=f"cuda:{ctx.device}")
    slots, stride, sizes, crcs, adlers = ctx.alloc_chunk_outputs(n, 65536, adler=True)
    slots.fill_(0xEE)
    ctx.deflate_chunks_primed(d_in, n, 65536, level, flush, slots, stride, sizes, crcs, adlers)
    torch.cuda.synchronize()
    nch = (n + 65535) // 65536
    u32 = lambda t: t.cpu().numpy().view(np.uint32)[:nch]
    return slots.cpu().numpy().reshape(-1, stride)[:nch], u32(sizes), u32(crcs), u32(adlers), stride


def assert_primed_parity(pkg, ctx, zo, data, flush=2, level=1):
    got, sizes, crcs, adlers, stride = gpu_primed(pkg, ctx, data, flush, level)
    exp, esizes, ecrcs, eadlers = zo.port_deflate_chunks_primed(data, 65536, level, flush, stride)
    bad = [i for i in range(len(esizes)) if sizes[i] != esizes[i] or not np.array_equal(got[i, : esizes[i]], exp[i, : esizes[i]])]
    assert not bad, f"level {level}: {len(bad)} of {len(esizes)} chunks differ, first {bad[0]}: gpu {sizes[bad[0]]} bytes, oracle {esizes[bad[0]]}"
    assert np.array_equal(crcs, ecrcs) and np.array_equal(adlers, eadlers)
    if zo.have_ref():
        ref, rsizes, _, _ = zo.ref_deflate_chunks_primed(data, 65536, level, flush, stride)
        assert np.array_equal(sizes, rsizes)
        assert all(np.array_equal(got[i, : rsizes[i]], ref[i, : rsizes[i]]) for i in range(len(rsizes)))
    return got, sizes


@pytest.mark.parametrize("level", LEVELS)
def test_primed_synthetic_mix_and_stream_validity(pkg, ctx, zo, level):
    data = synth(40 * 65536 + 4321, seed=61)
    got, sizes = assert_primed_parity(pkg, ctx, zo, data, 2, level)
    stream = b"".join(got[i, : sizes[i]].tobytes() for i in range(len(sizes))) + b"\x03\x00"
    assert pyzlib.decompress(stream, wbits=-15) == data.tobytes()                 # one dependent stream, inflated in order
    # data whose chunks resemble their predecessors: the dictionary recovers the ratio lost at the chunk joins
    rng = np.random.default_rng(8)
    words = rng.integers(97, 123, size=(64, 6), dtype=np.uint8)
    text = words[rng.integers(0, 64, size=60000)].reshape(-1)[:5 * 65536]
    _, psz = assert_primed_parity(pkg, ctx, zo, text, 2, level)
    assert int(psz.sum()) < int(zo.port_deflate_chunks(text, 65536, level, 2)[1].sum())


@pytest.mark.parametrize("level", LEVELS)
@pytest.mark.parametrize("flush", [2, 3, 4])
def test_primed_data_shapes(pkg, ctx, zo, flush, level):
    rng = np.random.default_rng(3)
    assert_primed_parity(pkg, ctx, zo, np.zeros(3 * 65536 + 5, dtype=np.uint8), flush, level)
    assert_primed_parity(pkg, ctx, zo, rng.integers(0, 256, size=2 * 65536 + 100, dtype=np.uint8), flush, level)
    assert_primed_parity(pkg, ctx, zo, rng.integers(0, 4, size=4 * 65536, dtype=np.uint8), flush, level)
    assert_primed_parity(pkg, ctx, zo, np.tile(np.arange(251, dtype=np.uint8), 1100)[:4 * 65536], flush, level)
    words = rng.integers(97, 123, size=(64, 6), dtype=np.uint8)
    assert_primed_parity(pkg, ctx, zo, words[rng.integers(0, 64, size=60000)].reshape(-1)[:5 * 65536], flush, level)


@pytest.mark.parametrize("level", LEVELS)
@pytest.mark.parametrize("tail", [0, 1, 2, 3, 4, 5, 100, 261, 262, 263, 300, 32767, 32768, 32769, 33000, 65274, 65275, 65279, 65535])
def test_primed_last_chunk_lengths(pkg, ctx, zo, tail, level):
    """The refill / slide schedule of fill_window (deflate.c:1272-1340) depends on the length of the primed chunk."""
    data = synth(2 * 65536 + tail, seed=tail)
    for flush in (2, 4):
        assert_primed_parity(pkg, ctx, zo, data, flush, level)
    assert_primed_parity(pkg, ctx, zo, data[: 65536 + tail], 2, level)
    assert_primed_parity(pkg, ctx, zo, data[:tail], 4, level)                     # a lone first chunk: no dictionary at all


def test_primed_slid_window_position_zero_alias(pkg, ctx, zo):
    rng = np.random.default_rng(1)
    for n in (65535, 65400, 65290, 65279):
        d = rng.integers(0, 256, size=65536 + n, dtype=np.uint8)
        c = d[65536:]
        c[32760:32780] = c[1000:1020]
        c[65274:65282] = c[32768:32776] if n >= 65282 else c[32768:32768 + n - 65274]
        assert_primed_parity(pkg, ctx, zo, d, 4)


def test_primed_golden_digests_of_the_unmodified_reference(pkg, ctx, golden):
    import zlib
    cases = golden("primed_digests.json")["cases"]                                # 13 level-1 cases + 60 at levels 2-6
    assert len(cases) >= 70 and sorted({c.get("level", 1) for c in cases}) == LEVELS
    for c in cases:
        data = synth(c["n"], seed=c["seed"])
        got, sizes, crcs, _, _ = gpu_primed(pkg, ctx, data, c["flush"], c.get("level", 1))
        assert [int(x) for x in sizes] == c["sizes"], c
        assert [int(zlib.crc32(got[i, : sizes[i]].tobytes())) for i in range(len(sizes))] == c["comp_crc32"], c


@pytest.mark.parametrize("level", [2, 6])
def test_primed_window_levels_64MiB_every_chunk_vs_reference(pkg, ctx, zo, level):
    """1 024 primed chunks of the synthetic mix at level 2 (deflate_fast) and 6 (deflate_medium with look-ahead): every chunk's bytes
    against the unmodified reference's own SetDictionary call sequence."""
    data = synth(64 << 20, seed=909 + level)
    got, sizes, _, _, stride = gpu_primed(pkg, ctx, data, 2, level)
    fn = zo.ref_deflate_chunks_primed if zo.have_ref() else zo.port_deflate_chunks_primed
    exp, esizes, _, _ = fn(data, 65536, level, 2, stride)
    bad, first = zo.compare_chunks(got, stride, sizes, exp, stride, esizes)
    assert bad == 0, (level, bad, first)
    stream = b"".join(got[i, : sizes[i]].tobytes() for i in range(len(sizes))) + b"\x03\x00"
    assert pyzlib.decompress(stream, wbits=-15) == data.tobytes()


def test_primed_256MiB_round_trip(pkg, ctx, zo):
    import torch
    n = 256 << 20
    data = synth(n, seed=2027)
    d_in = torch.from_numpy(data).to(f"cuda:{ctx.device}")
    slots, stride, sizes, crcs, _ = ctx.alloc_chunk_outputs(n, 65536)
    ctx.deflate_chunks_primed(d_in, n, 65536, 1, 2, slots, stride, sizes, crcs, None)
    torch.cuda.synchronize()
    nch = n // 65536
    sz = sizes.cpu().numpy().view(np.uint32)[:nch]
    host = slots.cpu().numpy().reshape(-1, stride)
    stream = b"".join(host[i, : sz[i]].tobytes() for i in range(nch)) + b"\x03\x00"
    assert pyzlib.decompress(stream, wbits=-15) == data.tobytes()
    for ci in np.random.default_rng(0).integers(1, nch, size=40):                 # sampled chunks against the oracle
        piece = data[(ci - 1) * 65536:(ci + 1) * 65536]
        exp, es, _, _ = zo.port_deflate_chunks_primed(piece, 65536, 1, 2, stride, nthreads=1)
        assert sz[ci] == es[1] and np.array_equal(host[ci, : es[1]], exp[1, : es[1]]), ci


def test_primed_stream_inflates_through_the_stream_path(pkg, ctx, zo):
    """A dependent stream has sync-flush markers but its segments reference their predecessors: the parallel marker split
    of zng_b200_inflate_stream_host must notice and give the exact in-order result."""
    rng = np.random.default_rng(8)
    words = rng.integers(97, 123, size=(64, 6), dtype=np.uint8)
    text = words[rng.integers(0, 64, size=90000)].reshape(-1)[:7 * 65536 + 999]
    for data in (text, synth(6 * 65536 + 5, seed=91)):
        got, sizes, _, _, _ = gpu_primed(pkg, ctx, data, 2)
        raw = b"".join(got[i, : sizes[i]].tobytes() for i in range(len(sizes))) + b"\x03\x00"
        assert pyzlib.decompress(raw, wbits=-15) == data.tobytes()
        src = np.frombuffer(raw + b"\0" * 8, dtype=np.uint8).copy()
        out = np.zeros(data.size + 16, dtype=np.uint8)
        status, out_len, in_used, check, detail = ctx.inflate_stream_host(src, len(raw), -15, out, out.size)
        assert status == 1, (status, detail)
        assert out_len == data.size and in_used == len(raw) and np.array_equal(out[:out_len], data)
