"""The oracle's primed level-1 restatement (pigz's dependent-chunk mode) against the committed digests of the
unmodified reference (tests/golden/primed_digests.json, written by tests/golden/make_golden.py) and, live, against
oracle/_ref: fresh zng_deflateInit2 + zng_deflateSetDictionary(32768 bytes in front of the chunk) + one zng_deflate(flush)."""
import zlib as pyzlib

import numpy as np
import pytest


def test_port_primed_matches_golden_digests(pkg, zo, golden):
    cases = golden("primed_digests.json")["cases"]
    assert len(cases) >= 10
    for c in cases:
        data = pkg.synth(c["n"], seed=c["seed"])
        out, sizes, _, _ = zo.port_deflate_chunks_primed(data, 65536, 1, c["flush"])
        assert [int(x) for x in sizes] == c["sizes"], c
        assert [int(pyzlib.crc32(out[i, : sizes[i]].tobytes())) for i in range(len(sizes))] == c["comp_crc32"], c


def test_port_primed_matches_reference_live(pkg, zo):
    if not zo.have_ref() or not hasattr(zo.ref(), "refdrv_deflate_chunks_primed"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    inputs = [pkg.synth(12 * 65536 + 777, seed=4), np.zeros(3 * 65536 + 5, dtype=np.uint8), rng.integers(0, 4, size=4 * 65536, dtype=np.uint8),
              np.tile(np.arange(251, dtype=np.uint8), 1100)[:4 * 65536]]
    inputs += [pkg.synth(2 * 65536 + t, seed=t) for t in (1, 2, 3, 4, 261, 262, 263, 32768, 32769, 65274, 65275, 65535)]
    for d in inputs:
        for flush in (2, 3, 4):
            a = zo.port_deflate_chunks_primed(d, 65536, 1, flush)
            b = zo.ref_deflate_chunks_primed(d, 65536, 1, flush)
            assert np.array_equal(a[1], b[1])
            assert all(np.array_equal(a[0][i, : a[1][i]], b[0][i, : b[1][i]]) for i in range(len(a[1])))
        st = b"".join(a[0][i, : a[1][i]].tobytes() for i in range(len(a[1])))     # flush 4: every chunk ends its own stream
    out, sizes, _, _ = zo.port_deflate_chunks_primed(inputs[0], 65536, 1, 2)
    stream = b"".join(out[i, : sizes[i]].tobytes() for i in range(len(sizes))) + b"\x03\x00"
    assert pyzlib.decompress(stream, wbits=-15) == inputs[0].tobytes()
