"""One test per BASELINE.json config at its full per-GPU size: EVERY unit (chunk / member / buffer) of the CUDA path's
output is compared with the unmodified reference (oracle/_ref, compiled from /root/reference by oracle/Makefile; the oracle
port when that library is absent) -- not a sample.  Inputs are the seeded synthetic workload (tests/synth.c).

configs[0]  level 1, 1 GiB = 16,384 chunks              test_config0_level1_1GiB_every_chunk
configs[1]  crc32 / adler32 + combine, 1 and 4 GiB      test_config1_checksums (16 GiB: slow)
configs[2]  levels 2 and 3, 4 GiB = 65,536 chunks       test_config2_level2_and_3_4GiB_every_chunk
configs[3]  1 Mi gzip members of 4 KiB                  test_config3_one_million_members
configs[4]  64 GiB on 8 GPUs = 8 GiB per GPU            test_config4_one_gpu_share_8GiB (slow); the 8-GPU run itself is
                                                        `bench.py --config 4 --gpus 8`, which compares every chunk on every rank
"""
import numpy as np
from synthdata import SEED, fill, synth
import pytest

pytestmark = pytest.mark.gpu

CHUNK = 65536
PIECE = 16384            # chunks per comparison piece (1 GiB)


def _deflate_every_chunk(pkg, ctx, zo, n, level, seed_offset=0):
    """GPU: the whole n-byte shard in one call; reference: piece by piece on every host core; compare every chunk."""
    import torch
    dev = torch.device("cuda", ctx.device)
    nch = n // CHUNK
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    fill(h_in.data_ptr(), n, SEED, seed_offset)
    d_in = h_in.to(dev)
    stride = pkg.deflate_bound(CHUNK)
    slots = torch.empty(nch * stride, dtype=torch.uint8, device=dev)
    slots.fill_(0xEE)
    sizes = torch.zeros(nch, dtype=torch.int32, device=dev); crcs = torch.zeros_like(sizes); adlers = torch.zeros_like(sizes)
    ctx.deflate_chunks(d_in, n, CHUNK, level, pkg.Z_FULL_FLUSH, slots, stride, sizes, crcs, adlers)
    res = torch.zeros(2, dtype=torch.int32, device=dev)
    ctx.crc32_fold(crcs, nch, CHUNK, n, 0, res[0:1])
    torch.cuda.synchronize()
    hs = sizes.cpu().numpy().view(np.uint32); hc = crcs.cpu().numpy().view(np.uint32); ha = adlers.cpu().numpy().view(np.uint32)
    hin = h_in.numpy()
    fold = 0
    comb = zo.ref().zng_crc32_combine if zo.have_ref() else zo.port().zo_crc32_combine
    for c0 in range(0, nch, PIECE):
        c1 = min(nch, c0 + PIECE)
        kind, exp, es, ec, ea = zo.best_deflate_chunks(hin[c0 * CHUNK:c1 * CHUNK], CHUNK, level, 3, stride)
        got = slots[c0 * stride:c1 * stride].cpu().numpy()
        bad, first = zo.compare_chunks(got, stride, hs[c0:c1], exp, stride, es)
        assert bad == 0, f"level {level}: {bad} of {c1 - c0} chunks differ from {kind}, first at chunk {c0 + first}"
        assert np.array_equal(hc[c0:c1], ec) and np.array_equal(ha[c0:c1], ea), f"level {level}: per-chunk crc32 / adler32 differ from {kind}"
        for k in range(c1 - c0):
            fold = comb(fold, int(ec[k]), CHUNK)
    assert (int(res[0].item()) & 0xffffffff) == fold                  # the trailer CRC of the assembled gzip stream
    return nch


def test_config0_level1_1GiB_every_chunk(pkg, ctx, zo):
    assert _deflate_every_chunk(pkg, ctx, zo, 1 << 30, 1) == 16384


@pytest.mark.parametrize("level", [2, 3])
def test_config2_level2_and_3_4GiB_every_chunk(pkg, ctx, zo, level):
    assert _deflate_every_chunk(pkg, ctx, zo, 4 << 30, level) == 65536


@pytest.mark.slow
def test_config4_one_gpu_share_8GiB(pkg, ctx, zo):
    """configs[4] gives each of 8 GPUs 8 GiB = 131,072 chunks of the 64 GiB stream; this is rank 3's share."""
    assert _deflate_every_chunk(pkg, ctx, zo, 8 << 30, 1, seed_offset=3 * (8 << 30)) == 131072


def _checksums(pkg, ctx, zo, n):
    import torch
    dev = torch.device("cuda", ctx.device)
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    fill(h_in.data_ptr(), n)
    d_in = h_in.to(dev)
    res = torch.zeros(4, dtype=torch.int32, device=dev)
    ctx.crc32(d_in, n, 0, res[0:1])
    ctx.adler32(d_in, n, 1, res[1:2])
    nt = (n + CHUNK - 1) // CHUNK
    tc = torch.zeros(nt, dtype=torch.int32, device=dev); ta = torch.zeros_like(tc)
    ctx.checksum_chunks(d_in, n, CHUNK, tc, ta)                       # the fused single pass + folds
    ctx.crc32_fold(tc, nt, CHUNK, n, 0, res[2:3])
    ctx.adler32_fold(ta, nt, CHUNK, n, 1, res[3:4])
    torch.cuda.synchronize()
    r = res.cpu().numpy().view(np.uint32)
    hin = h_in.numpy()
    ec, ea = (zo.ref_crc32(hin), zo.ref_adler32(hin)) if zo.have_ref() else (zo.port_crc32(hin), zo.port_adler32(hin))
    assert (int(r[0]), int(r[1])) == (ec, ea)
    assert (int(r[2]), int(r[3])) == (ec, ea)
    # sharded 2 / 4 / 8 ways + the reference's combine functions = the whole-buffer values (what the N-GPU run does)
    L = zo.ref() if zo.have_ref() else zo.port()
    ccomb = L.zng_crc32_combine if zo.have_ref() else L.zo_crc32_combine
    acomb = L.zng_adler32_combine if zo.have_ref() else L.zo_adler32_combine
    for g in (2, 8):
        per = n // g
        c, a = 0, 1
        pr = torch.zeros(2, dtype=torch.int32, device=dev)
        for k in range(g):
            ctx.crc32(d_in[k * per:(k + 1) * per], per, 0, pr[0:1])
            ctx.adler32(d_in[k * per:(k + 1) * per], per, 1, pr[1:2])
            torch.cuda.synchronize()
            v = pr.cpu().numpy().view(np.uint32)
            c = ccomb(c, int(v[0]), per); a = acomb(a, int(v[1]), per)
        assert (c, a) == (ec, ea)


@pytest.mark.parametrize("gib", [1, 4])
def test_config1_checksums(pkg, ctx, zo, gib):
    _checksums(pkg, ctx, zo, gib << 30)


@pytest.mark.slow
def test_config1_checksums_16GiB(pkg, ctx, zo):
    _checksums(pkg, ctx, zo, 16 << 30)


def test_config3_one_million_members(pkg, ctx, zo):
    """1 Mi independent 4 KiB gzip members written by the reference (zng_deflateInit2(1, 31) + Z_FINISH = minigzip -1 framing;
    the oracle port without oracle/_ref), inflated on the GPU; status / size / CRC-32 / bytes of every member against the
    reference's zng_inflate."""
    import torch
    nm = 1 << 20
    n = nm * 4096
    data = synth(n, seed=SEED)
    dev = torch.device("cuda", ctx.device)
    piece = 1 << 18
    for m0 in range(0, nm, piece):
        raw = data[m0 * 4096:(m0 + piece) * 4096]
        members, in_off = zo.gzip_members(raw, 4096, 1)
        out_off = np.arange(piece + 1, dtype=np.uint64) * 4096
        # reference inflate
        r_out = np.empty(piece * 4096, dtype=np.uint8)
        r_sizes = np.zeros(piece, dtype=np.uint32); r_crcs = np.zeros(piece, dtype=np.uint32); r_status = np.zeros(piece, dtype=np.int32)
        fn = zo.ref().refdrv_inflate_members if zo.have_ref() else zo.port().zo_inflate_members
        assert fn(members.ctypes.data, in_off.ctypes.data, piece, r_out.ctypes.data, out_off.ctypes.data, r_sizes.ctypes.data,
                  r_crcs.ctypes.data, r_status.ctypes.data, 16) == 0
        assert (r_status == 1).all() and np.array_equal(r_out, raw)
        # GPU inflate
        d_in = torch.from_numpy(members).to(dev)
        d_io = torch.from_numpy(in_off.astype(np.int64)).to(dev); d_oo = torch.from_numpy(out_off.astype(np.int64)).to(dev)
        d_out = torch.zeros(piece * 4096, dtype=torch.uint8, device=dev)
        g_sizes = torch.zeros(piece, dtype=torch.int32, device=dev); g_checks = torch.zeros_like(g_sizes); g_status = torch.zeros_like(g_sizes)
        ctx.inflate_members(d_in, d_io, piece, 31, d_out, d_oo, g_sizes, g_checks, g_status, None, None)
        torch.cuda.synchronize()
        assert np.array_equal(g_status.cpu().numpy(), r_status)
        assert np.array_equal(g_sizes.cpu().numpy().view(np.uint32), r_sizes)
        assert np.array_equal(g_checks.cpu().numpy().view(np.uint32), r_crcs)
        assert np.array_equal(d_out.cpu().numpy(), r_out)
