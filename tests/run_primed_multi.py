#!/usr/bin/env python3
"""N-GPU dependent (primed) stream == the 1-stream reference.  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_primed_multi.py [--level L] [--mib-per-gpu M]

Rank r owns a contiguous shard; zng_b200_halo_exchange (one ncclSend / ncclRecv pair per neighbour) brings the 32 KiB in front of
it from rank r-1, zng_b200_deflate_chunks_primed_at primes every chunk -- also the shard's first -- with the 32 KiB in front of it,
and zng_b200_stream_index_multi (the allgather of (size, crc32)) places the shards.  Every rank then checks its chunks byte for
byte against the reference's own call sequence (fresh stream + zng_deflateSetDictionary + zng_deflate per chunk) over the WHOLE
stream, i.e. against what one GPU, or the CPU, emits for the un-sharded input.  Test infrastructure (uses oracle/)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from __graft_entry__ import load_package, load_oracle
from synthdata import synth

ap = argparse.ArgumentParser(); ap.add_argument("--level", type=int, default=2); ap.add_argument("--mib-per-gpu", type=int, default=16)
args = ap.parse_args()
pkg = load_package(); zo = load_oracle()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ctx = pkg.Context(local); comm = pkg.Comm.from_torch(ctx)
n = args.mib_per_gpu << 20
nch = n // 65536
whole = synth(n * world, seed=4242)                       # every rank generates the whole stream (small): the checker needs it
mine = whole[rank * n:(rank + 1) * n]
buf = torch.zeros(32768 + n, dtype=torch.uint8, device=dev)          # [halo | shard], contiguous
buf[32768:].copy_(torch.from_numpy(mine.copy()))
d_halo, d_in = buf[:32768], buf[32768:]
comm.halo_exchange(d_in, n, d_halo)
stride = pkg.deflate_bound(65536)
slots = torch.empty(nch * stride, dtype=torch.uint8, device=dev)
sizes = torch.zeros(nch, dtype=torch.int32, device=dev); crcs = torch.zeros_like(sizes)
ctx.deflate_chunks_primed_at(d_in, n, 65536, args.level, pkg.Z_SYNC_FLUSH, rank > 0, slots, stride, sizes, crcs)
offs = torch.zeros(nch + 1, dtype=torch.int64, device=dev)
end, crc, tin = comm.stream_index(sizes, crcs, nch, 65536, n, 0, offs)
torch.cuda.synchronize()
fn = zo.ref_deflate_chunks_primed if zo.have_ref() else zo.port_deflate_chunks_primed
exp, es, ec, _ = fn(whole, 65536, args.level, 2, stride)
lo = rank * nch
bad, first = zo.compare_chunks(slots.cpu().numpy(), stride, sizes.cpu().numpy().view(np.uint32), exp[lo:lo + nch], stride, es[lo:lo + nch])
ok = bad == 0 and end == int(es.astype(np.int64).sum()) and tin == n * world
ok = ok and int(offs[0].item()) == int(es[:lo].astype(np.int64).sum())
import zlib
ok = ok and crc == zlib.crc32(whole.tobytes())
flag = torch.tensor([1 if ok else 0], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"primed level {args.level}, {world} rank(s) x {args.mib_per_gpu} MiB: " + ("every chunk of every rank equals the un-sharded reference stream; offsets, length and crc32 agree" if int(flag.item()) else f"MISMATCH (rank 0: {bad} chunks differ, first {first})"))
comm.close()
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
