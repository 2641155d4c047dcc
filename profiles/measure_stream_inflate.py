#!/usr/bin/env python3
"""Round trip of one 1 GiB gzip stream through the host library's own zng_deflate / zng_inflate (pinned buffers):
GB/s of uncompressed bytes, wall clock, best of 3."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package(); L = pkg.lib()
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
from __graft_entry__ import load_synth; load_synth().fill(h_in.data_ptr(), n)
for level in (1, 2):
    s = pkg.ZngStream()
    assert L.zng_deflateInit2(ctypes.byref(s), level, 8, 31, 8, 0) == 0
    cap = int(L.zng_deflateBound(ctypes.byref(s), n)) + 64
    comp = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    best_c = best_d = 1e9
    for it in range(4):
        L.zng_deflateReset(ctypes.byref(s))
        s.next_in = h_in.data_ptr(); s.avail_in = n; s.next_out = comp.data_ptr(); s.avail_out = cap
        t0 = time.perf_counter(); r = L.zng_deflate(ctypes.byref(s), 4); t1 = time.perf_counter()
        assert r == 1
        clen = int(s.total_out)
        if it: best_c = min(best_c, t1 - t0)
    L.zng_deflateEnd(ctypes.byref(s))
    back = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for it in range(4):
        d = pkg.ZngStream()
        assert L.zng_inflateInit2(ctypes.byref(d), 31) == 0
        d.next_in = comp.data_ptr(); d.avail_in = clen; d.next_out = back.data_ptr(); d.avail_out = n
        t0 = time.perf_counter(); r = L.zng_inflate(ctypes.byref(d), 4); t1 = time.perf_counter()
        assert r == 1 and d.total_out == n, (r, d.msg)
        L.zng_inflateEnd(ctypes.byref(d))
        if it: best_d = min(best_d, t1 - t0)
    assert torch.equal(back, h_in)
    print(f"level {level}: {n >> 20} MiB -> {clen} bytes (ratio {clen / n:.3f}); zng_deflate {n / best_c / 1e9:.2f} GB/s, zng_inflate {n / best_d / 1e9:.2f} GB/s (host to host, pinned)")
