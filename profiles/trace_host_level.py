import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from __graft_entry__ import load_package
from synthdata import fill
pkg = load_package()
level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = 1 << 30
ctx = pkg.Context(0)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); fill(h_in.data_ptr(), n)
cap = n // 65536 * pkg.deflate_bound(65536)
h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
for it in range(3):
    t0 = time.perf_counter(); r = ctx.deflate_host(h_in, n, 65536, level, False, h_out, cap); dt = time.perf_counter() - t0
    print(f"level {level} run {it}: {n / dt / 1e9:.2f} GB/s ({dt * 1e3:.0f} ms)", flush=True)
