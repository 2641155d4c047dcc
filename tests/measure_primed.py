#!/usr/bin/env python3
"""Device-resident throughput of zng_b200_deflate_chunks_primed (pigz's dependent mode) at levels 1-6 next to the independent-chunk
path, and the unmodified reference doing the same call sequence on the host cores.
python tests/measure_primed.py [MiB]   (lives under tests/: it times oracle/_ref as the CPU baseline)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from __graft_entry__ import load_package, load_oracle
from synthdata import synth
pkg = load_package(); zo = load_oracle()
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
ctx = pkg.Context(0)
data = synth(n)
d_in = torch.from_numpy(data).cuda()
slots, stride, sizes, crcs, _ = ctx.alloc_chunk_outputs(n, 65536)
for level in (1, 2, 3, 4, 5, 6):
    for name, fn in (("independent (Z_FULL_FLUSH)", lambda: ctx.deflate_chunks(d_in, n, 65536, level, 3, slots, stride, sizes, crcs, None)),
                     ("primed (dictionary = 32 KiB in front, Z_SYNC_FLUSH)", lambda: ctx.deflate_chunks_primed(d_in, n, 65536, level, 2, slots, stride, sizes, crcs, None))):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out = int(sizes.cpu().numpy().view(np.uint32)[: n // 65536].astype(np.int64).sum())
        print(f"level {level} GPU {name}: {n / ms / 1e6:.2f} GB/s ({ms:.1f} ms per {mib} MiB), ratio {out / n:.4f}", flush=True)
    if zo.have_ref():
        sample = data[: min(n, 128 << 20)]
        t0 = time.perf_counter(); r = zo.ref_deflate_chunks_primed(sample, 65536, level, 2); t1 = time.perf_counter()
        print(f"level {level} CPU reference primed: {sample.size / (t1 - t0) / 1e9:.2f} GB/s on {min(os.cpu_count() or 1, 32)} threads (128 MiB sample), ratio {int(r[1].astype(np.int64).sum()) / sample.size:.4f}", flush=True)
