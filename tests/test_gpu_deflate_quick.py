"""K1 parity: the sm_100a level-1 chunk compressor, called through the C ABI
(zng_b200_deflate_chunks), against the CPU oracle on the same inputs -- byte for byte -- plus the
committed digests of the unmodified reference and size-independent round-trip properties."""
import zlib as pyzlib   # independent inflater for round trips / byte-string digests only

import numpy as np
from synthdata import synth
import pytest

pytestmark = pytest.mark.gpu


def gpu_deflate(pkg, ctx, data, chunk=65536, level=1, flush=3, want_adler=True):
    import torch
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = data.size
    d_in = torch.from_numpy(data).to(f"cuda:{ctx.device}") if n else torch.zeros(16, dtype=torch.uint8, device=f"cuda:{ctx.device}")
    slots, stride, sizes, crcs, adlers = ctx.alloc_chunk_outputs(n, chunk, adler=want_adler)
    slots.fill_(0xEE)   # poison: bytes past `size` must never matter
    ctx.deflate_chunks(d_in, n, chunk, level, flush, slots, stride, sizes, crcs, adlers)
    torch.cuda.synchronize()
    nch = (n + chunk - 1) // chunk
    u32 = lambda t: t.cpu().numpy().view(np.uint32)[:nch]
    return slots.cpu().numpy().reshape(-1, stride)[:nch], u32(sizes), u32(crcs), (u32(adlers) if want_adler else None), stride


def explain_mismatch(pkg, ctx, zo, data, chunk, ci, level=1):
    """Token-level diff of chunk ci (the ZLIB_DEBUG trace analogue) for the assertion message."""
    import torch
    piece = np.ascontiguousarray(data[ci * chunk:(ci + 1) * chunk])
    exp = zo.port_tokens(piece, level)
    d_in = torch.from_numpy(piece).to(f"cuda:{ctx.device}")
    slots, stride, sizes, _, _ = ctx.alloc_chunk_outputs(piece.size, chunk)
    toks = torch.zeros(chunk + 8, dtype=torch.int32, device=d_in.device)
    ctx.deflate_chunks_trace(d_in, piece.size, chunk, level, 3, slots, stride, sizes, toks, chunk + 8)
    torch.cuda.synchronize()
    got = toks.cpu().numpy().view(np.uint32)
    end = np.nonzero(got == 0x40000000)[0]
    got = got[: end[0]] if len(end) else got
    k = 0
    while k < min(len(got), len(exp)) and got[k] == exp[k]:
        k += 1
    fmt = lambda t: f"match(len={(t >> 16) & 0x1ff},dist={t & 0xffff})" if t & 0x80000000 else f"lit({t})"
    return (f"chunk {ci}: {len(got)} tokens vs oracle {len(exp)}; first difference at token {k}: "
            f"gpu {fmt(int(got[k])) if k < len(got) else 'END'} vs oracle {fmt(int(exp[k])) if k < len(exp) else 'END'}")


def assert_parity(pkg, ctx, zo, data, chunk=65536, flush=3, level=1):
    got, sizes, crcs, adlers, stride = gpu_deflate(pkg, ctx, data, chunk, level, flush)
    exp, esizes, ecrcs, eadlers = zo.port_deflate_chunks(data, chunk, level, flush, stride)
    bad = [i for i in range(len(esizes)) if sizes[i] != esizes[i] or not np.array_equal(got[i, : esizes[i]], exp[i, : esizes[i]])]
    if bad:
        pytest.fail(f"{len(bad)} of {len(esizes)} chunks differ; " + explain_mismatch(pkg, ctx, zo, data, chunk, bad[0], level))
    assert np.array_equal(crcs, ecrcs)
    assert np.array_equal(adlers, eadlers)
    return got, sizes


def test_synthetic_mix_bit_exact(pkg, ctx, zo):
    data = synth(64 * 65536 + 4321, seed=101)       # every unit type several times + a ragged tail
    assert_parity(pkg, ctx, zo, data)


def test_each_unit_type(pkg, ctx, zo):
    base = synth(10 * 65536, seed=202)
    for u in range(10):
        assert_parity(pkg, ctx, zo, base[u * 65536:(u + 1) * 65536])


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 63, 64, 65, 257, 258, 259, 260, 262, 263, 4095, 4096, 65535])
def test_short_inputs(pkg, ctx, zo, n):
    data = synth(65536, seed=5)[:n]
    for flush in (3, 4):
        assert_parity(pkg, ctx, zo, data, 65536, flush)


def test_zeros_runs_and_random(pkg, ctx, zo):
    rng = np.random.default_rng(7)
    assert_parity(pkg, ctx, zo, np.zeros(3 * 65536 + 5, dtype=np.uint8))
    assert_parity(pkg, ctx, zo, rng.integers(0, 256, size=2 * 65536 + 100, dtype=np.uint8))
    assert_parity(pkg, ctx, zo, np.full(65536, 0xAB, dtype=np.uint8))
    # period patterns (dist-k matches clipped at 258), and a candidate-0 alias
    for period in (1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 255, 256, 257, 258, 259, 1000):
        pat = rng.integers(0, 256, size=period, dtype=np.uint8)
        assert_parity(pkg, ctx, zo, np.tile(pat, 65536 // period + 1)[:65536])
    two = rng.integers(0, 4, size=65536, dtype=np.uint8)       # tiny alphabet: many hash collisions in a window
    assert_parity(pkg, ctx, zo, two)


def test_matches_far_and_at_max_dist(pkg, ctx, zo):
    rng = np.random.default_rng(8)
    blk = rng.integers(0, 256, size=300, dtype=np.uint8)
    for gap in (32506 - 300, 32506 - 4, 32506 - 3, 32505, 32506, 32507, 32768, 40000):
        d = rng.integers(0, 256, size=65536, dtype=np.uint8)
        d[100:400] = blk
        d[100 + gap:400 + gap] = blk
        assert_parity(pkg, ctx, zo, d)


def test_small_chunks_and_finish_members(pkg, ctx, zo):
    data = synth(64 * 4096, seed=7)
    assert_parity(pkg, ctx, zo, data, 4096, 4)        # config-4 style members (one Z_FINISH block each)
    assert_parity(pkg, ctx, zo, synth(65536, seed=3)[:257 * 40], 257, 3)
    assert_parity(pkg, ctx, zo, synth(3 * 65536, seed=9), 1000, 3)


def test_golden_digests_of_the_unmodified_reference(pkg, ctx, golden):
    from test_oracle_deflate import golden_cases
    k = 0
    for c, data in golden_cases(pkg, golden):
        if c["level"] != 1:
            continue
        got, sizes, crcs, adlers, _ = gpu_deflate(pkg, ctx, data, c["chunk"], 1, c["flush"])
        assert [int(x) for x in sizes] == c["sizes"], c["name"]
        assert [int(pyzlib.crc32(got[i, : sizes[i]].tobytes())) for i in range(len(sizes))] == c["comp_crc32"], c["name"]
        assert [int(x) for x in crcs] == c["crc32"], c["name"]
        assert [int(x) for x in adlers] == c["adler32"], c["name"]
        k += 1
    assert k >= 30


def test_stream_assembly_round_trip_256MiB(pkg, ctx, zo):
    """Size-independent properties at scale: offsets are the prefix sum of sizes, the gathered
    stream inflates back to the input (independent inflater), and the fold of the per-chunk
    CRC-32s equals the CRC-32 of the whole buffer."""
    import torch
    n = 256 << 20
    data = synth(n, seed=303)
    dev = f"cuda:{ctx.device}"
    d_in = torch.from_numpy(data).to(dev)
    slots, stride, sizes, crcs, adlers = ctx.alloc_chunk_outputs(n)
    ctx.deflate_chunks(d_in, n, 65536, 1, 3, slots, stride, sizes, crcs, adlers)
    nch = n // 65536
    offsets = torch.zeros(nch + 1, dtype=torch.int64, device=dev)
    ctx.chunk_offsets(sizes, nch, 0, offsets)
    torch.cuda.synchronize()
    total = int(offsets[nch].item())
    hs = sizes.cpu().numpy().view(np.uint32).astype(np.int64)
    assert total == int(hs.sum())
    assert np.array_equal(offsets.cpu().numpy()[:-1], np.concatenate([[0], np.cumsum(hs)[:-1]]))
    packed = torch.empty(total + 2, dtype=torch.uint8, device=dev)
    ctx.gather_chunks(slots, stride, sizes, offsets, nch, packed)
    res = torch.zeros(2, dtype=torch.int32, device=dev)
    ctx.crc32_fold(crcs, nch, 65536, n, 0, res[0:1])
    ctx.adler32_fold(adlers, nch, 65536, n, 1, res[1:2])
    torch.cuda.synchronize()
    stream = packed.cpu().numpy()
    stream[total:total + 2] = (3, 0)                      # zng_deflate(Z_FINISH) with no input: "03 00"
    whole_crc = pyzlib.crc32(data.tobytes())
    r = res.cpu().numpy().view(np.uint32)
    assert int(r[0]) == whole_crc
    assert int(r[1]) == pyzlib.adler32(data.tobytes())
    if zo.have_ref():
        code, out_len, crc = zo.ref_inflate_stream(stream, -15, expect=data)
        assert code == 1 and out_len == n                 # Z_STREAM_END, bytes identical (checked inside)
    else:
        assert pyzlib.decompress(stream.tobytes(), wbits=-15) == data.tobytes()
    # spot-check bit-exactness on a sample of chunks against the oracle
    pick = np.random.default_rng(1).choice(nch, size=96, replace=False)
    hslots = slots.cpu().numpy().reshape(-1, stride)
    for ci in pick:
        exp, es, _, _ = zo.port_deflate_chunks(data[ci * 65536:(ci + 1) * 65536], 65536, 1, 3, stride, nthreads=1)
        assert es[0] == hs[ci] and np.array_equal(hslots[ci, : es[0]], exp[0, : es[0]]), f"chunk {ci}"


def test_argument_errors(pkg, ctx):
    import torch
    d = torch.zeros(65536, dtype=torch.uint8, device=f"cuda:{ctx.device}")
    slots, stride, sizes, crcs, adlers = ctx.alloc_chunk_outputs(65536)
    for kwargs in (dict(level=0), dict(level=7), dict(level=9), dict(flush=0), dict(flush=5), dict(chunk=0), dict(chunk=65537), dict(stride=stride - 16), dict(stride=stride + 8)):
        a = dict(chunk=65536, level=1, flush=3, stride=stride)
        a.update(kwargs)
        with pytest.raises(pkg.ZngB200Error) as ei:
            ctx.deflate_chunks(d, 65536, a["chunk"], a["level"], a["flush"], slots, a["stride"], sizes, crcs, adlers)
        assert ei.value.code in (pkg.Z_STREAM_ERROR, pkg.Z_BUF_ERROR)


def test_slid_window_position_zero_alias(pkg, ctx, zo):
    """A chunk of 65275..65535 bytes is slid with the parser at 65274; slide_hash zeroes the entries below 32768 and
    deflate_quick, which has no hash_head != 0 test, then tries window position 0 (= byte 32768) at distance MAX_DIST."""
    rng = np.random.default_rng(1)
    for n in (65535, 65400, 65290, 65279, 65536):
        d = rng.integers(0, 256, size=n, dtype=np.uint8)
        d[32760:32780] = d[1000:1020]          # a match covers 32768, so that position is not inserted
        d[65274:65282] = d[32768:32776] if n >= 65282 else d[32768:32768 + n - 65274]
        for flush in (3, 4):
            assert_parity(pkg, ctx, zo, d, 65536, flush)


@pytest.mark.parametrize("knobs", [("warp", "", "", ""), ("cta", "3", "1", "10"), ("cta", "2", "2", "11"), ("cta", "4", "1", "8"), ("cta", "1", "4", "6")])
def test_both_level1_parsers_are_bit_exact(pkg, zo, knobs, monkeypatch):
    """The two level-1 parsers forced by ZNG_B200_K1: "warp" = K1a v3 (csrc/deflate_quick.cu, one warp per chain), "cta" = K1a v7
    (csrc/deflate_quick_cta.cu: producer warps one window ahead of a walker warp; several shapes).  Same
    inputs as the shipped parser's tests -- the synthetic mix with a ragged tail, every short length class, repetitive data, the
    post-slide position-0 alias -- compared byte for byte with the oracle, plus 64 MiB against the unmodified reference."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    monkeypatch.setenv("ZNG_B200_K1", knobs[0])         # default is "auto": v7 up to 8192 chunks per launch, the warp-per-chain parser beyond
    if knobs[1]:
        monkeypatch.setenv("ZNG_B200_K1_WARPS", knobs[1]); monkeypatch.setenv("ZNG_B200_K1_BPW", knobs[2]); monkeypatch.setenv("ZNG_B200_K1_CHAINS", knobs[3])
    c = pkg.Context(0)                                   # the knobs are read when a context is created
    try:
        rng = np.random.default_rng(11)
        inputs = [synth(64 * 65536 + 4321, seed=101), np.zeros(3 * 65536 + 17, dtype=np.uint8), rng.integers(0, 4, size=5 * 65536, dtype=np.uint8),
                  np.tile(rng.integers(0, 256, size=37, dtype=np.uint8), 8000)[:4 * 65536 + 5], rng.integers(0, 256, size=2 * 65536 + 1, dtype=np.uint8)]
        inputs += [synth(65536, seed=5)[:n] for n in (0, 1, 3, 4, 5, 31, 32, 33, 95, 96, 97, 127, 128, 129, 255, 256, 257, 4096, 65275, 65279, 65535)]
        for data in inputs:
            for flush in (3, 4):
                assert_parity(pkg, c, zo, data, 65536, flush)
        assert_parity(pkg, c, zo, synth(64 * 4096, seed=7), 4096, 4)
        d = rng.integers(0, 256, size=65400, dtype=np.uint8)                     # slid window, position-0 alias (see below)
        d[32760:32780] = d[1000:1020]; d[65274:65282] = d[32768:32776]
        assert_parity(pkg, c, zo, d, 65536, 3)
        big = synth(64 << 20, seed=808)
        got, sizes, crcs, _, stride = gpu_deflate(pkg, c, big, want_adler=False)
        _, exp, es, ec, _ = zo.best_deflate_chunks(big, 65536, 1, 3, stride)
        bad, first = zo.compare_chunks(got, stride, sizes, exp, stride, es)
        assert bad == 0 and np.array_equal(crcs, ec), (knobs, bad, first)
    finally:
        c.close()
