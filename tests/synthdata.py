"""The synthetic mixed text/binary workload of BASELINE.json -- TEST / BENCH INFRASTRUCTURE.

tests/synth.c is compiled into tests/libzng_synth.so by `make -C oracle synth` (__graft_entry__.build()); it is not part
of the product library.  Every 64 KiB unit depends on (seed, unit index) alone, so ranks / threads can fill disjoint
ranges and agree on the bytes.
"""
from __future__ import annotations

import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

SEED = 0x9E3779B97F4A7C15
UNIT = 65536
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libzng_synth.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle synth`")
        L = ctypes.CDLL(LIB_PATH)
        L.zng_synth_fill.restype = ctypes.c_int
        L.zng_synth_fill.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64]
        _lib = L
    return _lib


def fill(ptr: int, n: int, seed: int = SEED, offset: int = 0, threads: int | None = None) -> None:
    """Fill n bytes at address `ptr` with the stream bytes [offset, offset + n); offset must be a multiple of 65536."""
    L = lib()
    threads = threads or min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 4)
    units = (n + UNIT - 1) // UNIT
    if threads <= 1 or units < 64:
        r = L.zng_synth_fill(ptr, n, seed, offset)
        if r != 0:
            raise RuntimeError(f"zng_synth_fill failed: {r}")
        return
    per = (units + threads - 1) // threads * UNIT

    def part(k):
        o = k * per
        return L.zng_synth_fill(ptr + o, min(per, n - o), seed, offset + o) if o < n else 0
    with ThreadPoolExecutor(threads) as ex:
        if any(r != 0 for r in ex.map(part, range(threads))):
            raise RuntimeError("zng_synth_fill failed")


def synth(n: int, seed: int = SEED, offset: int = 0) -> np.ndarray:
    """n bytes of the synthetic workload as a host numpy array."""
    buf = np.empty(n, dtype=np.uint8)
    if n:
        fill(buf.ctypes.data, n, seed, offset)
    return buf
