"""The N > 1 path on CPU: two `gloo` ranks shard a synthetic buffer by contiguous chunk ranges, exchange the per-chunk
(size, crc32) pairs with one allgather, scan + fold them (zlib-ng_b200/stream.py) and rank 0 assembles ONE gzip stream
that an independent inflater accepts.  The chunk compressor itself is a CUDA kernel; here the oracle stands in for it
(test only) so that the host-side sharding / collective / framing logic is exercised without a GPU."""
import os
import sys
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, level, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_oracle, load_package
    pkg = load_package()
    zo = load_oracle()
    import importlib.util
    spec = importlib.util.spec_from_file_location("zstream", os.path.join(ROOT, "zlib-ng_b200", "stream.py"))
    st = importlib.util.module_from_spec(spec); spec.loader.exec_module(st)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        data = synth(n, seed=77)
        nch = (n + st.CHUNK - 1) // st.CHUNK
        lo, hi = st.shard_range(nch, rank, world)
        mine = data[lo * st.CHUNK: min(hi * st.CHUNK, n)]
        out, sizes, crcs, _ = zo.port_deflate_chunks(mine, st.CHUNK, level, 3)
        if lo > 0 and level == 2 and mine.size % st.CHUNK:
            # a short last chunk sees the previous chunk's window bytes: compress it in stream context
            o2, s2, c2, _ = zo.port_deflate_chunks(data[(hi - 2) * st.CHUNK: n], st.CHUNK, level, 3)
            out[-1, :] = 0; out[-1, : s2[1]] = o2[1, : s2[1]]; sizes[-1] = s2[1]
        all_sizes, all_crcs = st.allgather_pairs(torch.from_numpy(sizes.astype(np.int32)), torch.from_numpy(crcs.astype(np.int64).astype(np.int32)))
        offsets, crc = st.scan_and_fold(pkg.lib(), all_sizes.numpy(), all_crcs.numpy(), n)
        payload = b"".join(out[i, : sizes[i]].tobytes() for i in range(len(sizes)))
        # this rank's bytes land at offsets[lo] .. offsets[hi] of the final stream
        assert len(payload) == int(offsets[hi] - offsets[lo])
        gathered = [None] * world
        dist.all_gather_object(gathered, (int(offsets[lo]), payload))
        if rank == 0:
            total = int(offsets[-1])
            stream = bytearray(total + 2 + 8)
            stream[:10] = st.gzip_header(level)
            for off, pl in gathered:
                stream[off: off + len(pl)] = pl
            stream[total: total + 2] = st.FINISH_EMPTY
            stream[total + 2:] = st.gzip_trailer(crc, n)
            ok = pyzlib.decompress(bytes(stream), wbits=31) == data.tobytes() and crc == pyzlib.crc32(data.tobytes())
            if zo.have_ref():       # the unmodified reference inflates the assembled stream to the original bytes
                code, out_len, rcrc = zo.ref_inflate_stream(bytes(stream), 31, expect=data)
                ok = ok and code == 1 and out_len == n
            q.put((ok, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,level", [(9 * 65536, 1), (7 * 65536 + 1234, 1), (6 * 65536, 2)])
def test_two_ranks_assemble_one_gzip_stream(n, level):
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, n, level, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    ok, lo, hi = q.get(timeout=10)
    assert ok


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("zstream", os.path.join(ROOT, "zlib-ng_b200", "stream.py"))
    st = importlib.util.module_from_spec(spec); spec.loader.exec_module(st)
    for nch in (0, 1, 2, 7, 8, 16384, 1048576 + 3):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = st.shard_range(nch, r, world)
                assert lo == prev and hi >= lo and hi - lo in (nch // world, nch // world + 1)
                prev = hi
            assert prev == nch
