"""The oracle's primed restatements (level 1: quick_parse_primed; levels 2-6: the window engine) (pigz's dependent-chunk mode) against the committed digests of the
unmodified reference (tests/golden/primed_digests.json, written by tests/golden/make_golden.py) and, live, against
oracle/_ref: fresh zng_deflateInit2 + zng_deflateSetDictionary(32768 bytes in front of the chunk) + one zng_deflate(flush)."""
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest


def test_port_primed_matches_golden_digests(pkg, zo, golden):
    cases = golden("primed_digests.json")["cases"]
    assert len(cases) >= 10
    for c in cases:
        data = synth(c["n"], seed=c["seed"])
        out, sizes, _, _ = zo.port_deflate_chunks_primed(data, 65536, c.get("level", 1), c["flush"])
        assert [int(x) for x in sizes] == c["sizes"], c
        assert [int(pyzlib.crc32(out[i, : sizes[i]].tobytes())) for i in range(len(sizes))] == c["comp_crc32"], c


def test_port_primed_matches_reference_live(pkg, zo):
    if not zo.have_ref() or not hasattr(zo.ref(), "refdrv_deflate_chunks_primed"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    inputs = [synth(12 * 65536 + 777, seed=4), np.zeros(3 * 65536 + 5, dtype=np.uint8), rng.integers(0, 4, size=4 * 65536, dtype=np.uint8),
              np.tile(np.arange(251, dtype=np.uint8), 1100)[:4 * 65536]]
    inputs += [synth(2 * 65536 + t, seed=t) for t in (1, 2, 3, 4, 261, 262, 263, 32768, 32769, 65274, 65275, 65535)]
    for d in inputs:
        for level in (1, 2, 3, 4, 5, 6):
            for flush in ((2, 3, 4) if level == 1 else (2, 4)):
                a = zo.port_deflate_chunks_primed(d, 65536, level, flush)
                b = zo.ref_deflate_chunks_primed(d, 65536, level, flush)
                assert np.array_equal(a[1], b[1]), (level, flush)
                assert all(np.array_equal(a[0][i, : a[1][i]], b[0][i, : b[1][i]]) for i in range(len(a[1]))), (level, flush)
        st = b"".join(a[0][i, : a[1][i]].tobytes() for i in range(len(a[1])))     # flush 4: every chunk ends its own stream
    out, sizes, _, _ = zo.port_deflate_chunks_primed(inputs[0], 65536, 1, 2)
    stream = b"".join(out[i, : sizes[i]].tobytes() for i in range(len(sizes))) + b"\x03\x00"
    assert pyzlib.decompress(stream, wbits=-15) == inputs[0].tobytes()


def test_window_engine_agrees_with_the_chunk_coordinate_restatement(pkg, zo):
    """Levels 2-6 have two independent restatements in the oracle: chunk coordinates with a modelled tail slide
    (fast_parse / medium_parse) and the window engine that keeps the reference's own window / head / prev state.
    On fresh streams they must agree byte for byte (and both are pinned to the reference elsewhere)."""
    rng = np.random.default_rng(9)
    words = rng.integers(97, 123, size=(64, 6), dtype=np.uint8)
    inputs = [synth(6 * 65536, seed=14), rng.integers(0, 4, size=2 * 65536, dtype=np.uint8), words[rng.integers(0, 64, size=30000)].reshape(-1)[:2 * 65536],
              synth(65536, seed=2)[:65300], synth(65536, seed=3)[:65275], synth(65536, seed=4)[:300], np.zeros(65536, dtype=np.uint8)]
    for d in inputs:
        for level in (2, 3, 4, 5, 6):
            a = zo.port_deflate_chunks_fresh_window(d, 65536, level, 4)
            for i in range(len(a[1])):
                b = zo.port_deflate_chunks(d[i * 65536:(i + 1) * 65536], 65536, level, 4, nthreads=1)
                assert a[1][i] == b[1][0] and np.array_equal(a[0][i, : a[1][i]], b[0][0, : b[1][0]]), (level, i)


def test_primed_token_trace_is_consistent(pkg, zo):
    """zo_deflate_tokens_primed (what a kernel under construction is diffed against): the tokens cover the chunk exactly, every
    match copies bytes that are really there -- reaching into the dictionary where the data repeats across the join."""
    rng = np.random.default_rng(4)
    words = rng.integers(97, 123, size=(64, 6), dtype=np.uint8)
    text = words[rng.integers(0, 64, size=40000)].reshape(-1)[:32768 + 65536]
    for data in (text, synth(32768 + 65536, seed=6), synth(32768 + 1000, seed=7)):
        chunk = data[32768:]
        for level in (1, 2, 6):
            toks = zo.port_tokens_primed(data, level)
            pos, into_dict = 32768, 0
            for t in toks:
                t = int(t)
                if t & 0x80000000:
                    ln, dist = (t >> 16) & 0x1ff, t & 0xffff
                    assert 3 <= ln <= 258 and 1 <= dist <= 32768 - 262 and pos - dist >= 0
                    assert np.array_equal(data[pos:pos + ln], np.array([data[pos - dist + (k % dist)] for k in range(ln)], dtype=np.uint8)) or \
                        bytes(data[pos - dist:pos - dist + ln]) == bytes(data[pos:pos + ln])
                    into_dict += pos - dist < 32768
                    pos += ln
                else:
                    assert t == int(data[pos])
                    pos += 1
            assert pos == data.size, (level, pos, data.size)
            if data is text:
                assert into_dict > 0
