"""Host-side assembly of ONE gzip stream from chunk ranges compressed by several ranks.

Rank r of G owns the contiguous chunk range shard_range(nchunks, r, G).  The only exchange is an allgather of the
per-chunk (compressed size, crc32) pairs -- 8 bytes per chunk -- after which every rank knows
  * the byte offset of each of its chunks in the final stream (exclusive prefix sum; chunk outputs are byte aligned
    because each ends with the empty stored block of Z_FULL_FLUSH, deflate.c:1064-1065), and
  * the CRC-32 of the whole input (fold of crc32_combine, crc32_braid_comb.c:16-24) for the gzip trailer
    (deflate.c:1091-1096: CRC32 LE, ISIZE = total_in mod 2^32).
Stream = header(10) | chunks | "03 00" (zng_deflate(Z_FINISH) with no input) | crc32 | isize.
The collective runs on whatever torch.distributed backend the process group has (nccl on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import struct

import numpy as np

CHUNK = 65536
GZIP_HEADER_LEN = 10
FINISH_EMPTY = b"\x03\x00"


def shard_range(nchunks: int, rank: int, world: int):
    """Contiguous chunk range [lo, hi) of `rank` (sizes differ by at most one chunk)."""
    base, rem = divmod(nchunks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gzip_header(level: int) -> bytes:
    """deflate.c:902-921 with no gz_header: 1f 8b 08 00 <mtime 0> XFL OS; XFL 4 for level 1, OS_CODE 3 (unix)."""
    return bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4 if level < 2 else 0, 3])


def gzip_trailer(crc: int, total_in: int) -> bytes:
    return struct.pack("<II", crc & 0xffffffff, total_in & 0xffffffff)


def allgather_pairs(sizes, crcs, world_counts=None):
    """All-gather this rank's per-chunk (size, crc32) pairs; returns (all_sizes, all_crcs) in global chunk order.
    sizes / crcs: 1-D int32 torch tensors on the process group's device.  Ranks may own different counts."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    mine = torch.stack([sizes.view(-1), crcs.view(-1)]).to(torch.int32)
    n_local = torch.tensor([mine.shape[1]], dtype=torch.int64, device=mine.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    width = max(counts) if counts else 0
    padded = torch.zeros(2, width, dtype=torch.int32, device=mine.device)
    padded[:, : mine.shape[1]] = mine
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    all_sizes = torch.cat([o[0, :c] for o, c in zip(out, counts)])
    all_crcs = torch.cat([o[1, :c] for o, c in zip(out, counts)])
    return all_sizes, all_crcs


def scan_and_fold(lib, all_sizes: np.ndarray, all_crcs: np.ndarray, n_total: int, chunk: int = CHUNK, base: int = GZIP_HEADER_LEN):
    """offsets[i] = base + sum_{j<i} size_j (nchunks + 1 entries) and the crc32 of the whole input by folding
    crc32_combine over the chunks with the library's own combine functions (host arithmetic on two words)."""
    sizes = np.asarray(all_sizes).astype(np.uint32).astype(np.int64)
    crcs = np.asarray(all_crcs).astype(np.uint32)
    offsets = np.zeros(sizes.size + 1, dtype=np.int64)
    offsets[0] = base
    np.cumsum(sizes, out=offsets[1:])
    offsets[1:] += base
    crc = 0
    op_full = lib.zng_crc32_combine_gen(chunk)
    for i in range(crcs.size):
        ln = min(chunk, n_total - i * chunk)
        crc = lib.zng_crc32_combine_op(crc, int(crcs[i]), op_full) if ln == chunk else lib.zng_crc32_combine(crc, int(crcs[i]), ln)
    return offsets, crc
