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
