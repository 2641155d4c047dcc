"""K2 parity: the sm_100a level-2 .. level-6 chunk compressors (deflate_fast / deflate_medium: hash chains +
longest_match, dynamic / static / stored blocks from the trees.c block writer), called through the C ABI, against the
CPU oracle byte for byte, the committed digests of the unmodified reference, and round trips through an independent
inflater.  Every test runs for every level."""
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

from test_gpu_deflate_quick import assert_parity, gpu_deflate

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[2, 3, 4, 5, 6], ids=["level2", "level3", "level4", "level5", "level6"])
def level(request):
    return request.param


def test_synthetic_mix_bit_exact(pkg, ctx, zo, level):
    data = synth(64 * 65536 + 4321, seed=101)
    got, sizes = assert_parity(pkg, ctx, zo, data, level=level)
    # every chunk inflates on its own (Z_FULL_FLUSH boundaries) with an independent inflater
    for i in (0, 1, 2, 3, 17, 63, 64):
        piece = data[i * 65536:(i + 1) * 65536].tobytes()
        d = pyzlib.decompressobj(-15)
        assert d.decompress(got[i, : sizes[i]].tobytes()) == piece


def test_each_unit_type(pkg, ctx, zo, level):
    base = synth(10 * 65536, seed=202)
    for u in range(10):
        assert_parity(pkg, ctx, zo, base[u * 65536:(u + 1) * 65536], level=level)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 63, 64, 65, 257, 258, 259, 260, 262, 263, 4095, 4096, 65535])
def test_short_inputs(pkg, ctx, zo, n, level):
    data = synth(65536, seed=5)[:n]
    for flush in (3, 4):
        assert_parity(pkg, ctx, zo, data, 65536, flush, level=level)


def test_zeros_runs_and_random(pkg, ctx, zo, level):
    rng = np.random.default_rng(7)
    assert_parity(pkg, ctx, zo, np.zeros(3 * 65536 + 5, dtype=np.uint8), level=level)
    assert_parity(pkg, ctx, zo, rng.integers(0, 256, size=2 * 65536 + 100, dtype=np.uint8), level=level)      # stored blocks
    assert_parity(pkg, ctx, zo, np.full(65536, 0xAB, dtype=np.uint8), level=level)
    for period in (1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 255, 256, 257, 258, 259, 1000):
        pat = rng.integers(0, 256, size=period, dtype=np.uint8)
        assert_parity(pkg, ctx, zo, np.tile(pat, 65536 // period + 1)[:65536], level=level)
    for alphabet in (2, 3, 4, 16):                     # tiny alphabets: long hash chains, many collisions inside a window
        assert_parity(pkg, ctx, zo, rng.integers(0, alphabet, size=65536, dtype=np.uint8), level=level)


def test_chain_walk_and_nice_match(pkg, ctx, zo, level):
    """Several candidates with growing common prefixes (3..7 bytes, then >= 8): exercises best_len updates, the
    chain limit of 4 and the nice_match stop (match_tpl.h)."""
    rng = np.random.default_rng(11)
    d = rng.integers(0, 256, size=65536, dtype=np.uint8)
    key = rng.integers(0, 256, size=40, dtype=np.uint8)
    pos = 1000
    for k in (4, 4, 5, 6, 7, 4, 9, 5, 30, 4, 4, 4, 4, 4, 12):
        d[pos:pos + k] = key[:k]
        d[pos + k] = (int(key[k]) + 1 + k) & 0xff
        pos += 97 + 13 * k
    d[pos:pos + 40] = key
    assert_parity(pkg, ctx, zo, d, level=level)
    words = [rng.integers(97, 123, size=int(rng.integers(3, 9)), dtype=np.uint8) for _ in range(50)]
    text = np.concatenate([np.concatenate([words[int(rng.integers(0, 50))], np.array([32], dtype=np.uint8)]) for _ in range(14000)])[:65536]
    assert_parity(pkg, ctx, zo, text, level=level)


def test_virtual_bytes_past_the_end(pkg, ctx, zo, level):
    """longest_match does not clamp nice_match to lookahead, so the last match of a chunk depends
    on what lies past the data: chunk[32768+k] for a full chunk (after the tail slide), the previous chunk's bytes
    for a short chunk on the same stream, zeros for a short first chunk."""
    rng = np.random.default_rng(12)
    for trial in range(12):
        d = rng.integers(0, 256, size=2 * 65536 + 30000, dtype=np.uint8)
        for c in range(3):
            end = min((c + 1) * 65536, d.size)
            tail = rng.integers(0, 256, size=int(rng.integers(4, 9)), dtype=np.uint8)
            base = c * 65536
            for back in (3000, 9000, 15531, 25531):          # the same tail earlier in the chunk, followed by different bytes
                d[end - back - tail.size:end - back] = tail
            d[end - tail.size:end] = tail
            # make one of the earlier copies continue with the virtual bytes
            if c < 2:
                d[end - 9000:end - 9000 + 6] = d[base + 32768:base + 32774]
        assert_parity(pkg, ctx, zo, d, level=level)
        assert_parity(pkg, ctx, zo, d[:30000], level=level)      # short first chunk: zeros past the data


def test_matches_far_and_at_max_dist(pkg, ctx, zo, level):
    rng = np.random.default_rng(8)
    blk = rng.integers(0, 256, size=300, dtype=np.uint8)
    for gap in (32506 - 300, 32506 - 4, 32506 - 3, 32505, 32506, 32507, 32768, 40000):
        d = rng.integers(0, 256, size=65536, dtype=np.uint8)
        d[100:400] = blk
        d[100 + gap:400 + gap] = blk
        assert_parity(pkg, ctx, zo, d, level=level)


def test_block_splits_at_16383_symbols(pkg, ctx, zo, level):
    """A chunk of pure literals fills blocks of exactly 16383 symbols (deflate_fast.c:93-94); 65536 = 4 * 16383 + 4,
    and 3 * 16383 / 4 * 16383 exercise the 'block full exactly at the end of input' paths for both flush modes."""
    rng = np.random.default_rng(13)
    lit = rng.permutation(np.arange(65536, dtype=np.uint32) * 2654435761 % 251).astype(np.uint8)
    for n in (16382, 16383, 16384, 2 * 16383, 3 * 16383 + 1, 4 * 16383, 65536):
        for flush in (3, 4):
            assert_parity(pkg, ctx, zo, rng.integers(0, 256, size=n, dtype=np.uint8), 65536, flush, level=level)
            assert_parity(pkg, ctx, zo, lit[:n], 65536, flush, level=level)


def test_lookahead_hand_over_and_fizzle(pkg, ctx, zo, level):
    """Many short chunks: every chunk ends in the zone where deflate_medium stops looking ahead (deflate_medium.c:164-172,
    234); low-entropy and periodic data make adjacent matches that fizzle_matches (:84-144) shifts."""
    rng = np.random.default_rng(77)
    data = synth(6 * 65536, seed=31)
    for chunk in (553, 554, 555, 600, 777, 1024, 2500, 8191):
        assert_parity(pkg, ctx, zo, data[:200 * chunk if 200 * chunk < data.size else data.size], chunk, 3, level=level)
    low = rng.integers(0, 3, size=4 * 65536, dtype=np.uint8)
    assert_parity(pkg, ctx, zo, low, level=level)
    assert_parity(pkg, ctx, zo, low[:150000], 3000, 4, level=level)
    words = rng.integers(97, 101, size=(16, 5), dtype=np.uint8)
    text = words[rng.integers(0, 16, size=60000)].reshape(-1)
    assert_parity(pkg, ctx, zo, text[:4 * 65536], level=level)
    assert_parity(pkg, ctx, zo, text[:100 * 1234], 1234, 3, level=level)
    for n in (65536 - 262, 65536 - 263, 65536 - 261, 65274, 65275, 65273, 65535, 65534):
        assert_parity(pkg, ctx, zo, np.concatenate([text[:65536], text[1000:1000 + n]]), level=level)
        assert_parity(pkg, ctx, zo, low[:n], 65536, 4, level=level)


def test_small_chunks_and_finish_members(pkg, ctx, zo, level):
    data = synth(64 * 4096, seed=7)
    assert_parity(pkg, ctx, zo, data, 4096, 4, level=level)
    assert_parity(pkg, ctx, zo, synth(65536, seed=3)[:257 * 40], 257, 3, level=level)
    assert_parity(pkg, ctx, zo, synth(3 * 65536, seed=9), 1000, 3, level=level)


def test_golden_digests_of_the_unmodified_reference(pkg, ctx, golden, level):
    from test_oracle_deflate import golden_cases
    k = 0
    for c, data in golden_cases(pkg, golden):
        if c["level"] != level:
            continue
        got, sizes, crcs, adlers, _ = gpu_deflate(pkg, ctx, data, c["chunk"], level, c["flush"])
        assert [int(x) for x in sizes] == c["sizes"], c["name"]
        assert [int(pyzlib.crc32(got[i, : sizes[i]].tobytes())) for i in range(len(sizes))] == c["comp_crc32"], c["name"]
        assert [int(x) for x in crcs] == c["crc32"], c["name"]
        k += 1
    assert k >= 30


def test_host_path_level2_stream(pkg, ctx, zo, level):
    from test_gpu_host_path import oracle_stream
    for n, final in ((5 * 65536 + 99, True), (5 * 65536 + 99, False), ((70 << 20) + 12345, True)):
        data = synth(n, seed=n % 1000 + 1)
        cap = pkg.deflate_bound(65536) * ((n + 65535) // 65536 + 1)
        out = np.empty(cap, dtype=np.uint8)
        out_len, crc, adler = ctx.deflate_host(data, n, 65536, level, final, out, cap)
        if zo.have_ref() and final:
            exp = zo.ref_deflate_stream(data, 65536, level, -15).tobytes()
        else:
            exp = oracle_stream(zo, data, 65536, level, final)
        assert out_len == len(exp) and out[:out_len].tobytes() == exp
        assert crc == pyzlib.crc32(data.tobytes())
        if final:
            assert pyzlib.decompress(out[:out_len].tobytes(), wbits=-15) == data.tobytes()


def test_256MiB_round_trip_and_sampled_parity(pkg, ctx, zo, level):
    import torch
    n = 256 << 20
    data = synth(n, seed=404)
    dev = f"cuda:{ctx.device}"
    d_in = torch.from_numpy(data).to(dev)
    slots, stride, sizes, crcs, adlers = ctx.alloc_chunk_outputs(n)
    ctx.deflate_chunks(d_in, n, 65536, level, 3, slots, stride, sizes, crcs, None)
    nch = n // 65536
    offsets = torch.zeros(nch + 1, dtype=torch.int64, device=dev)
    ctx.chunk_offsets(sizes, nch, 0, offsets)
    torch.cuda.synchronize()
    total = int(offsets[nch].item())
    packed = torch.empty(total + 2, dtype=torch.uint8, device=dev)
    ctx.gather_chunks(slots, stride, sizes, offsets, nch, packed)
    torch.cuda.synchronize()
    stream = packed.cpu().numpy()
    stream[total:total + 2] = (3, 0)
    assert pyzlib.decompress(stream.tobytes(), wbits=-15) == data.tobytes()
    hs = sizes.cpu().numpy().view(np.uint32)
    hslots = slots.cpu().numpy().reshape(-1, stride)
    pick = np.random.default_rng(2).choice(nch, size=96, replace=False)
    for ci in pick:
        exp, es, _, _ = zo.port_deflate_chunks(data[ci * 65536:(ci + 1) * 65536], 65536, level, 3, stride, nthreads=1)
        assert es[0] == hs[ci] and np.array_equal(hslots[ci, : es[0]], exp[0, : es[0]]), f"chunk {ci}"


def test_k2c_cta_per_chunk_parser_is_bit_exact():
    """K2c (deflate_fast.cu: medium_cta_kernel, opt-in with ZNG_B200_K2C=1): inserter / searcher / resolver warps on one chunk, the
    prev[] ring in shared memory.  The switch is read once per process, so the level-5 and level-6 cases of this file run again in
    a child process with it set -- same oracle, same digests of the unmodified reference."""
    import os, subprocess, sys
    env = dict(os.environ, ZNG_B200_K2C="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_deflate_fast.py"), "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "(level5 or level6) and not 256MiB"], env=env, capture_output=True, text=True, timeout=900, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-500:]
