"""The oracle's deflate restatement (oracle/zo_deflate.c) against
  * the committed digests of the unmodified reference's output (tests/golden/deflate_digests.json,
    produced by tests/golden/make_golden.py from oracle/_ref), and
  * the unmodified reference itself, live, when oracle/_ref is present."""
import zlib as pyzlib   # only to digest byte strings (crc32) and to round-trip through an independent inflater

import numpy as np
from synthdata import synth
import pytest


def _case_input(pkg, name):
    """Rebuild the input of a golden case from its name (mirrors make_golden.deflate_digests)."""
    rng = np.random.default_rng(20261018)
    base = name.rsplit("_l", 1)[0]
    if base == "synth_2MiB":
        return synth(32 * 65536)
    if base == "synth_ragged":
        return synth(10 * 65536 + 777, seed=12345)
    if base == "synth_finish":
        return synth(4 * 65536 + 4097, seed=99)
    if base == "synth_4k_members":
        return synth(64 * 4096, seed=7)
    if base == "zeros":
        return np.zeros(3 * 65536 + 5, dtype=np.uint8)
    if base == "random":
        lvl = int(name.rsplit("_l", 1)[1])
        # the levels drew one after the other from the same generator: replay the draws
        draw = None
        for _ in range(lvl):
            draw = rng.integers(0, 256, size=2 * 65536 + 100, dtype=np.uint8)
        return draw
    if base == "tiny_sizes":
        return synth(65536, seed=3)[:257 * 40]
    if base.startswith("short_"):
        n = int(base.split("_")[1])
        return synth(65536, seed=5)[:n]
    raise KeyError(name)


def golden_cases(pkg, golden):
    for c in golden("deflate_digests.json")["cases"]:
        yield c, _case_input(pkg, c["name"])


def test_port_matches_golden_digests(pkg, zo, golden):
    ncase = 0
    for c, data in golden_cases(pkg, golden):
        assert data.size == c["n"], c["name"]
        out, sizes, crcs, adlers = zo.port_deflate_chunks(data, c["chunk"], c["level"], c["flush"])
        assert [int(x) for x in sizes] == c["sizes"], c["name"]
        assert [int(pyzlib.crc32(out[i, : sizes[i]].tobytes())) for i in range(len(sizes))] == c["comp_crc32"], c["name"]
        assert [int(x) for x in crcs] == c["crc32"], c["name"]
        assert [int(x) for x in adlers] == c["adler32"], c["name"]
        ncase += 1
    assert ncase >= 200


def test_port_matches_reference_live(pkg, zo):
    if not zo.have_ref():
        pytest.skip("oracle/_ref not built")
    for seed, n, chunk, flush in ((11, 48 * 65536, 65536, 3), (12, 5 * 65536 + 31000, 65536, 4), (13, 200 * 1000, 1000, 3)):
        data = synth(n, seed=seed)
        for level in (1, 2, 3, 4, 5, 6):
            a = zo.port_deflate_chunks(data, chunk, level, flush)
            # one thread: a short last chunk reads the stale window of the chunk its stream compressed before;
            # with a pool that would be whichever chunk the worker happened to take, not the previous one
            b = zo.ref_deflate_chunks(data, chunk, level, flush, nthreads=1)
            assert np.array_equal(a[1], b[1]), (seed, level)
            for i in range(len(a[1])):
                assert np.array_equal(a[0][i, : a[1][i]], b[0][i, : b[1][i]]), (seed, level, i)
            assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


def test_port_slid_window_position_zero_alias(pkg, zo):
    """Level 1 after the tail slide of a 65275..65535-byte chunk: a zeroed hash slot names window position 0 (deflate.c:1285-1299,
    deflate_quick.c:88-92 has no hash_head != 0 test)."""
    if not zo.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1)
    for n in (65535, 65400, 65290, 65536):
        d = rng.integers(0, 256, size=n, dtype=np.uint8)
        d[32760:32780] = d[1000:1020]
        d[65274:65282] = d[32768:32776]
        for level in (1, 2, 3, 6):
            a = zo.port_deflate_chunks(d, 65536, level, 4, nthreads=1)
            b = zo.ref_deflate_chunks(d, 65536, level, 4, nthreads=1)
            assert a[1][0] == b[1][0] and np.array_equal(a[0][0, : a[1][0]], b[0][0, : b[1][0]]), (n, level)


def test_port_output_inflates(pkg, zo):
    # independent check of validity: CPython's zlib inflates the concatenated chunk stream
    data = synth(6 * 65536 + 999, seed=21)
    for level in (1, 2, 3, 4, 5, 6):
        out, sizes, _, _ = zo.port_deflate_chunks(data, 65536, level, 3)
        stream = b"".join(out[i, : sizes[i]].tobytes() for i in range(len(sizes))) + b"\x03\x00"
        assert pyzlib.decompress(stream, wbits=-15) == data.tobytes()


def test_synth_is_deterministic_and_shardable(pkg):
    a = synth(20 * 65536)
    b = synth(8 * 65536, offset=12 * 65536)
    assert np.array_equal(a[12 * 65536:], b)
    assert not np.array_equal(synth(65536, seed=1), synth(65536, seed=2))


def _fuzz_input(rng, n):
    """Structured random bytes: literals from a small alphabet, runs, copies of earlier substrings at assorted distances."""
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    alphabet = rng.integers(0, 256, size=int(rng.integers(2, 40)), dtype=np.uint8)
    while pos < n:
        kind = rng.integers(0, 4)
        ln = int(min(n - pos, rng.integers(1, 400 if kind else 40)))
        if kind == 0 or pos == 0:
            out[pos:pos + ln] = alphabet[rng.integers(0, alphabet.size, size=ln)]
        elif kind == 1:
            out[pos:pos + ln] = out[pos - 1]
        else:
            dist = int(min(pos, rng.choice([1, 2, 3, 4, 7, 64, 258, 1000, 4096, 32506, 32507, 32768, 40000])))
            for k in range(ln):
                out[pos + k] = out[pos + k - dist]
        pos += ln
    return out


def test_port_matches_reference_on_structured_random_inputs(pkg, zo):
    """Deterministic fuzz: 120 structured inputs x levels 1-6, fresh stream per chunk (Z_FINISH), plus the same data as
    primed chunks -- the oracle must reproduce the unmodified reference byte for byte."""
    if not zo.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(2026)
    sizes = [int(x) for x in rng.integers(0, 3000, size=80)] + [int(x) for x in rng.integers(60000, 65537, size=30)] + [65536] * 10
    for n in sizes:
        d = _fuzz_input(rng, n)
        for level in (1, 2, 3, 4, 5, 6):
            a = zo.port_deflate_chunks(d, 65536, level, 4, nthreads=1)
            b = zo.ref_deflate_chunks(d, 65536, level, 4, nthreads=1)
            assert len(a[1]) == len(b[1]) and all(a[1][i] == b[1][i] and np.array_equal(a[0][i, : a[1][i]], b[0][i, : b[1][i]]) for i in range(len(a[1]))), (n, level)
    for n in sizes[-25:]:
        d = _fuzz_input(rng, 65536 + n)
        for level in (1, 2, 3, 4, 5, 6):
            a = zo.port_deflate_chunks_primed(d, 65536, level, 2, nthreads=1)
            b = zo.ref_deflate_chunks_primed(d, 65536, level, 2, nthreads=1)
            assert all(a[1][i] == b[1][i] and np.array_equal(a[0][i, : a[1][i]], b[0][i, : b[1][i]]) for i in range(len(a[1]))), ("primed", n, level)
