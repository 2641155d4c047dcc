#!/usr/bin/env python3
"""Latency of zng_b200_deflate_chunks (level 1, device resident, one call + synchronise) for small batches: the shipped
warp-per-chain parser against the CTA-per-chain parser (K1a v7).  python profiles/latency_small_batches.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from __graft_entry__ import load_package
from synthdata import synth
pkg = load_package()
data = synth(4096 * 65536)
d_all = torch.from_numpy(data).cuda()
for knob in (("warp", "", ""), ("cta", "3", "1"), ("cta", "4", "1"), ("cta", "2", "2")):
    os.environ["ZNG_B200_K1"] = knob[0]
    if knob[1]: os.environ["ZNG_B200_K1_WARPS"] = knob[1]; os.environ["ZNG_B200_K1_BPW"] = knob[2]
    ctx = pkg.Context(0)
    row = []
    for nch in (1, 16, 148, 592, 1480, 4096):
        n = nch * 65536
        slots, stride, sizes, crcs, _ = ctx.alloc_chunk_outputs(n, 65536)
        best = 1e9
        for _ in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.deflate_chunks(d_all[:n], n, 65536, 1, 3, slots, stride, sizes, crcs, None)
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        row.append(f"{nch}: {best * 1e3:.2f} ms ({n / best / 1e9:.2f} GB/s)")
    print(f"K1={knob[0]} producers={knob[1] or '-'} blocks={knob[2] or '-'}:  " + "   ".join(row), flush=True)
    ctx.close()
