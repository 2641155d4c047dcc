"""K3 parity: CRC-32 / Adler-32 kernels and the combine folds, through the C ABI, against the
reference's known-answer vectors and the CPU oracle."""
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

pytestmark = pytest.mark.gpu


def test_crc32_kats_through_host_entry(pkg, ctx, golden):
    for v in golden("kat_crc32.json")["vectors"]:
        if v["data_hex"] is None or v["len"] == 0:
            continue                                   # NULL / empty: host-library argument handling (test_host_api)
        d = np.frombuffer(bytes.fromhex(v["data_hex"]), dtype=np.uint8)
        assert ctx.crc32_host(d, d.size, v["init"]) == v["expect"], v


def test_adler32_kats_through_host_entry(pkg, ctx, golden):
    for v in golden("kat_adler32.json")["vectors"]:
        if v["data_hex"] is None:
            continue
        d = np.frombuffer(bytes.fromhex(v["data_hex"]), dtype=np.uint8)
        assert ctx.adler32_host(d, d.size, v["init"]) == v["expect"], v


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 255, 256, 257, 511, 512, 513, 5552, 65535, 65536, 65537, 3 * 65536 + 17, (8 << 20) + 3])
def test_flat_checksums_vs_oracle(pkg, ctx, zo, n):
    import torch
    rng = np.random.default_rng(n)
    data = rng.integers(0, 256, size=n, dtype=np.uint8)
    d = torch.from_numpy(data).to(f"cuda:{ctx.device}") if n else torch.zeros(16, dtype=torch.uint8, device=f"cuda:{ctx.device}")
    res = torch.zeros(2, dtype=torch.int32, device=d.device)
    for ci, ai in ((0, 1), (0xdeadbeef, 0x12345678 % (65521 << 16 | 65520))):
        ai = ((ai >> 16) % 65521) << 16 | (ai & 0xffff) % 65521
        ctx.crc32(d, n, ci, res[0:1])
        ctx.adler32(d, n, ai, res[1:2])
        torch.cuda.synchronize()
        r = res.cpu().numpy().view(np.uint32)
        assert int(r[0]) == zo.port_crc32(data, ci)
        assert int(r[1]) == zo.port_adler32(data, ai)


def test_unaligned_device_pointer(pkg, ctx, zo):
    import torch
    data = synth(4 * 65536, seed=77)
    d = torch.from_numpy(data).to(f"cuda:{ctx.device}")
    res = torch.zeros(1, dtype=torch.int32, device=d.device)
    for off in (1, 2, 3, 5, 15):
        ctx.crc32(d[off:], data.size - off, 0, res)
        torch.cuda.synchronize()
        assert int(res.cpu().numpy().view(np.uint32)[0]) == zo.port_crc32(data[off:])


def test_chunk_checksums_and_folds(pkg, ctx, zo):
    import torch
    n = 37 * 65536 + 1234
    data = synth(n, seed=88)
    dev = f"cuda:{ctx.device}"
    d = torch.from_numpy(data).to(dev)
    for tile in (65536, 4096, 1000):
        nt = (n + tile - 1) // tile
        crcs = torch.zeros(nt, dtype=torch.int32, device=dev)
        adlers = torch.zeros(nt, dtype=torch.int32, device=dev)
        ctx.checksum_chunks(d, n, tile, crcs, adlers)
        res = torch.zeros(2, dtype=torch.int32, device=dev)
        ctx.crc32_fold(crcs, nt, tile, n, 0, res[0:1])
        ctx.adler32_fold(adlers, nt, tile, n, 1, res[1:2])
        torch.cuda.synchronize()
        hc = crcs.cpu().numpy().view(np.uint32); ha = adlers.cpu().numpy().view(np.uint32)
        for i in (0, 1, nt // 2, nt - 1):
            piece = data[i * tile:(i + 1) * tile]
            assert int(hc[i]) == zo.port_crc32(piece) and int(ha[i]) == zo.port_adler32(piece)
        r = res.cpu().numpy().view(np.uint32)
        assert int(r[0]) == zo.port_crc32(data) and int(r[1]) == zo.port_adler32(data)


def test_crc_of_1GiB_property(pkg, ctx):
    """Checksum of checksums at BASELINE size: the device CRC-32/Adler-32 of a 1 GiB buffer equals
    the fold of its two halves and an independent host computation."""
    import torch
    n = 1 << 30
    data = synth(n, seed=404)
    d = torch.from_numpy(data).to(f"cuda:{ctx.device}")
    res = torch.zeros(4, dtype=torch.int32, device=d.device)
    ctx.crc32(d, n, 0, res[0:1])
    ctx.adler32(d, n, 1, res[1:2])
    ctx.crc32(d[: n // 2], n // 2, 0, res[2:3])
    torch.cuda.synchronize()
    half = int(res.cpu().numpy().view(np.uint32)[2])
    ctx.crc32(d[n // 2:], n // 2, half, res[3:4])       # chaining through `init` == combine
    torch.cuda.synchronize()
    r = res.cpu().numpy().view(np.uint32)
    assert int(r[0]) == pyzlib.crc32(data) == int(r[3])
    assert int(r[1]) == pyzlib.adler32(data)


def test_flat_checksums_at_4GiB_equal_the_combine_of_their_parts(pkg, ctx):
    """BASELINE configs[1] at full size: zng_crc32 / zng_adler32 of a 4 GiB device buffer must equal the
    crc32_combine / adler32_combine fold of its four 1 GiB quarters (a checksum of checksums), and the first quarter
    must agree with an independent CPU implementation."""
    import torch
    L = pkg.lib()
    gib = 1 << 30
    n = 4 * gib
    dev = f"cuda:{ctx.device}"
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = np.empty(gib, dtype=np.uint8)
    for q in range(4):
        h = synth(gib, offset=q * gib)
        d[q * gib:(q + 1) * gib] = torch.from_numpy(h).to(dev)
        if q == 0:
            first = (pyzlib.crc32(h.tobytes()), pyzlib.adler32(h.tobytes()))
    res = torch.zeros(10, dtype=torch.int32, device=dev)
    ctx.crc32(d, n, 0, res[0:1])
    ctx.adler32(d, n, 1, res[1:2])
    for q in range(4):
        ctx.crc32(d[q * gib:(q + 1) * gib], gib, 0, res[2 + 2 * q: 3 + 2 * q])
        ctx.adler32(d[q * gib:(q + 1) * gib], gib, 1, res[3 + 2 * q: 4 + 2 * q])
    torch.cuda.synchronize()
    r = [int(x) for x in res.cpu().numpy().view(np.uint32)]
    assert (r[2], r[3]) == first
    crc, adler = r[2], r[3]
    for q in range(1, 4):
        crc = L.zng_crc32_combine(crc, r[2 + 2 * q], gib)
        adler = L.zng_adler32_combine(adler, r[3 + 2 * q], gib)
    assert (r[0], r[1]) == (crc, adler)
    # running values continue across calls (what zng_crc32(crc, buf, len) promises)
    ctx.crc32(d[gib:], n - gib, r[2], res[0:1])
    ctx.adler32(d[gib:], n - gib, r[3], res[1:2])
    torch.cuda.synchronize()
    r2 = [int(x) for x in res.cpu().numpy().view(np.uint32)[:2]]
    assert r2 == [crc, adler]
