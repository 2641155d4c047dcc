"""The multi-GPU stream assembly behind the C ABI (csrc/multi.cu: zng_b200_comm_*, zng_b200_stream_index_multi,
zng_b200_gzip_multi) on ONE rank -- the communicator of size 1 runs the same scan / fold / pack code without NCCL traffic.
The N > 1 collective itself is exercised by `bench.py --gpus N` (every chunk of every rank against the reference, plus the
folded CRC-32 and the total length across ranks) and, for the host-side sharding logic, by tests/test_multi_rank.py (gloo)."""
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 65536, 37 * 65536 + 4321])
@pytest.mark.parametrize("level", [1, 2])
def test_gzip_multi_single_rank_is_a_valid_gzip_file_equal_to_the_reference_stream(pkg, ctx, zo, n, level):
    import torch
    dev = torch.device("cuda", ctx.device)
    data = synth(max(n, 1), seed=n % 1000 + 5)[:n]
    comm = pkg.Comm(ctx, 1, 0, None)
    try:
        nch = (n + 65535) // 65536
        stride = pkg.deflate_bound(65536)
        d_in = torch.from_numpy(data.copy()).to(dev) if n else torch.zeros(16, dtype=torch.uint8, device=dev)
        slots = torch.empty(max(nch, 1) * stride, dtype=torch.uint8, device=dev)
        sizes = torch.zeros(max(nch, 1), dtype=torch.int32, device=dev); crcs = torch.zeros_like(sizes)
        offs = torch.zeros(nch + 1, dtype=torch.int64, device=dev)
        packed = torch.zeros(max(nch, 1) * stride, dtype=torch.uint8, device=dev)
        my_off, my_bytes, file_bytes, hdr, trl = comm.gzip_multi(d_in, n, level, slots, stride, sizes, crcs, offs, packed, packed.numel())
        torch.cuda.synchronize()
        assert my_off == 10 and file_bytes == 10 + my_bytes + 10
        blob = hdr + packed[:my_bytes].cpu().numpy().tobytes() + trl
        assert len(blob) == file_bytes
        assert pyzlib.decompress(blob, wbits=31) == data.tobytes()
        if zo.have_ref():
            code, out_len, crc = zo.ref_inflate_stream(np.frombuffer(blob, dtype=np.uint8), 31, expect=data if n else None)
            assert code == 1 and out_len == n
        # the body is the concatenation of the reference's chunks
        if n:
            _, exp, es, ec, _ = zo.best_deflate_chunks(data, 65536, level, 3, stride)
            bad, _ = zo.compare_chunks(packed[:my_bytes].cpu().numpy(), 0, es, exp, stride, es)
            assert bad == 0
        # the index alone
        end, crc, tin = comm.stream_index(sizes, crcs, nch, 65536, n, 10, offs)
        assert end == 10 + my_bytes and tin == n and crc == pyzlib.crc32(data.tobytes())
        assert np.array_equal(offs.cpu().numpy()[:-1], 10 + np.concatenate([[0], np.cumsum(sizes.cpu().numpy()[:nch].astype(np.int64))])[:-1]) or nch == 0
    finally:
        comm.close()


def test_comm_argument_errors(pkg, ctx):
    with pytest.raises(pkg.ZngB200Error):
        pkg.Comm(ctx, 2, 5, b"\0" * 128)          # rank out of range
    with pytest.raises(pkg.ZngB200Error):
        pkg.Comm(ctx, 0, 0, None)


@pytest.mark.parametrize("level", [1, 2, 5])
def test_primed_shard_with_halo_equals_the_unsharded_stream(pkg, ctx, zo, level):
    """A rank's shard of a dependent stream (zng_b200_deflate_chunks_primed_at, have_halo): the 32768 bytes in front of the buffer prime
    its first chunk, so the shard's chunks equal chunks [k, k + m) of the un-sharded stream.  (The halo's trip between ranks is
    tests/run_primed_multi.py under torchrun.)"""
    import torch
    dev = torch.device("cuda", ctx.device)
    whole = synth(12 * 65536, seed=31 + level)
    k, m = 5, 7
    buf = torch.from_numpy(whole[k * 65536 - 32768:(k + m) * 65536].copy()).to(dev)
    d_in = buf[32768:]
    stride = pkg.deflate_bound(65536)
    slots = torch.empty(m * stride, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(m, dtype=torch.int32, device=dev); crcs = torch.zeros_like(sizes)
    ctx.deflate_chunks_primed_at(d_in, m * 65536, 65536, level, pkg.Z_SYNC_FLUSH, True, slots, stride, sizes, crcs)
    torch.cuda.synchronize()
    fn = zo.ref_deflate_chunks_primed if zo.have_ref() else zo.port_deflate_chunks_primed
    exp, es, _, _ = fn(whole, 65536, level, 2, stride)
    bad, first = zo.compare_chunks(slots.cpu().numpy(), stride, sizes.cpu().numpy().view(np.uint32), exp[k:k + m], stride, es[k:k + m])
    assert bad == 0, (level, bad, first)
