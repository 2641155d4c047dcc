#!/usr/bin/env python3
"""pigz_multi_gpu.py -- BASELINE configs[4]: one gzip stream from N GPUs.

    torchrun --nnodes=1 --nproc-per-node N zlib-ng_b200/tools/pigz_multi_gpu.py --gib-per-gpu 8 --out /tmp/big.gz --verify

Rank r of N owns the contiguous chunk range [r, r+1) * n/N of a synthetic stream, compresses it on its GPU as
independent 64 KiB level-1 chunks with Z_FULL_FLUSH ends, and the only collective is an NCCL allgather of the per-chunk
(compressed size, crc32) pairs (8 bytes per chunk).  Every rank then scans the sizes (its bytes' offset in the file) and
folds the CRCs (crc32_combine), packs its chunks and writes them at its offset; rank 0 adds the gzip header, the "03 00"
terminator and the trailer.  --verify reads the finished file back with CPython's zlib (an independent inflater); the check
against the unmodified reference's zng_inflate lives with the tests: tests/verify_gzip_with_reference.py FILE.
"""
import argparse
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib-per-gpu", type=float, default=1.0)
    ap.add_argument("--out", default="/tmp/zng_b200_multi.gz")
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--verify", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package, load_synth
    pkg = load_package()
    sd = load_synth()                  # the synthetic input of BASELINE configs[4] (tests/synthdata.py)
    spec = importlib.util.spec_from_file_location("zstream", os.path.join(ROOT, "zlib-ng_b200", "stream.py"))
    st = importlib.util.module_from_spec(spec); spec.loader.exec_module(st)

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    CH = st.CHUNK
    n_rank = int(args.gib_per_gpu * (1 << 30)) // CH * CH
    n_total = n_rank * world
    nch = n_rank // CH
    ctx = pkg.Context(local)
    stride = pkg.deflate_bound(CH)

    # this rank's shard, produced piecewise on the host and copied up
    d_in = torch.empty(n_rank, dtype=torch.uint8, device=dev)
    piece = 256 << 20
    h = torch.empty(min(piece, n_rank), dtype=torch.uint8, pin_memory=True)
    for o in range(0, n_rank, piece):
        k = min(piece, n_rank - o)
        sd.fill(h.data_ptr(), k, sd.SEED, rank * n_rank + o)
        d_in[o:o + k].copy_(h[:k]); torch.cuda.synchronize()

    slots = torch.empty(nch * stride, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(nch, dtype=torch.int32, device=dev); crcs = torch.zeros(nch, dtype=torch.int32, device=dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.deflate_chunks(d_in, n_rank, CH, args.level, pkg.Z_FULL_FLUSH, slots, stride, sizes, crcs, None)
    if world > 1:
        all_sizes, all_crcs = st.allgather_pairs(sizes, crcs)          # the one collective: 8 bytes per chunk
    else:
        all_sizes, all_crcs = sizes, crcs
    nall = nch * world
    offsets_all = torch.zeros(nall + 1, dtype=torch.int64, device=dev)
    res = torch.zeros(2, dtype=torch.int32, device=dev)
    ctx.chunk_offsets(all_sizes.contiguous(), nall, st.GZIP_HEADER_LEN, offsets_all)
    ctx.crc32_fold(all_crcs.contiguous(), nall, CH, n_total, 0, res[0:1])
    local_off = torch.zeros(nch + 1, dtype=torch.int64, device=dev)
    ctx.chunk_offsets(sizes, nch, 0, local_off)
    torch.cuda.synchronize()
    my_bytes = int(local_off[nch].item())
    packed = torch.empty(my_bytes + 16, dtype=torch.uint8, device=dev)
    ctx.gather_chunks(slots, stride, sizes, local_off, nch, packed)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    my_off = int(offsets_all[rank * nch].item()); end_off = int(offsets_all[nall].item())
    crc = int(res[0].item()) & 0xffffffff

    # every rank writes its bytes at its offset of the one file
    if rank == 0:
        with open(args.out, "wb") as f:
            f.truncate(end_off + 2 + 8)
    if world > 1:
        dist.barrier()
    host = packed[:my_bytes].cpu().numpy()
    fd = os.open(args.out, os.O_WRONLY)
    pos = 0
    while pos < my_bytes:
        pos += os.pwrite(fd, host[pos:pos + (64 << 20)].tobytes(), my_off + pos)
    if rank == 0:
        os.pwrite(fd, st.gzip_header(args.level), 0)
        os.pwrite(fd, st.FINISH_EMPTY + st.gzip_trailer(crc, n_total), end_off)
    os.close(fd)
    if world > 1:
        dist.barrier()
    t2 = time.perf_counter()
    tt = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        total = end_off + 10
        print(f"{world} GPU(s): {n_total / (1 << 30):.1f} GiB -> {total} bytes (ratio {total / n_total:.3f}), crc32 {crc:08x}; "
              f"compress + allgather + scan + fold + pack: {float(tt.item()) * 1e3:.1f} ms = {n_total / float(tt.item()) / 1e9:.1f} GB/s; "
              f"file written in {t2 - t1:.1f} s", flush=True)
        if args.verify:
            tv = time.perf_counter()
            import zlib
            d = zlib.decompressobj(31); out_len = 0; rcrc = 0
            with open(args.out, "rb") as f:
                while True:
                    b = f.read(64 << 20)
                    if not b:
                        break
                    o = d.decompress(b); out_len += len(o); rcrc = zlib.crc32(o, rcrc)
            code = 1 if d.eof else -5
            who = "CPython zlib"
            ok = code == 1 and out_len == n_total and (rcrc & 0xffffffff) == crc
            print(f"verify with {who}: code {code}, {out_len} bytes, crc32 {rcrc & 0xffffffff:08x} -> {'OK' if ok else 'MISMATCH'} ({time.perf_counter() - tv:.1f} s)", flush=True)
            if not ok:
                sys.exit(1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
