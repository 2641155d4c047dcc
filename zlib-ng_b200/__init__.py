"""zlib-ng_b200 -- Python (ctypes) harness over libzng_b200.so.

The product is the shared library: hand-written sm_100a kernels behind the C ABI of
``include/zng_b200.h`` plus the C11 host library that mirrors zlib-ng's own API
(``include/zlib-ng-b200.h``).  This module only binds those C symbols for the tests and
``bench.py``; PyTorch is used for device memory, streams and ``torch.distributed`` plumbing.
There is no fallback: importing works without a GPU (so the symbol table can be checked), but
every compute entry point fails loudly when the library or the device is missing.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_int, c_int32, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzng_b200.so")

Z_OK, Z_STREAM_END, Z_NEED_DICT = 0, 1, 2
Z_STREAM_ERROR, Z_DATA_ERROR, Z_MEM_ERROR, Z_BUF_ERROR, Z_VERSION_ERROR = -2, -3, -4, -5, -6
Z_NO_FLUSH, Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_FINISH = 0, 2, 3, 4
CHUNK_MAX = 65536

_lib = None


class ZngB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"zng_b200 error {code}: {msg}")
        self.code = code


def lib() -> ctypes.CDLL:
    """Load libzng_b200.so (built in-tree by ``make -C zlib-ng_b200``); no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `make -C {_HERE}` (or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32p, u64p = c_void_p, c_void_p, c_void_p
    sigs = {
        "zng_b200_device_count": (c_int, []),
        "zng_b200_ctx_create": (c_int, [POINTER(c_void_p), c_int]),
        "zng_b200_ctx_destroy": (None, [vp]),
        "zng_b200_ctx_device": (c_int, [vp]),
        "zng_b200_ctx_sm_count": (c_int, [vp]),
        "zng_b200_last_error": (c_char_p, [vp]),
        "zng_b200_sync": (c_int, [vp, vp]),
        "zng_b200_host_alloc": (c_void_p, [c_size_t]),
        "zng_b200_host_free": (None, [vp]),
        "zng_b200_deflate_bound": (c_size_t, [c_size_t]),
        "zng_b200_deflate_chunks": (c_int, [vp, vp, c_size_t, c_uint32, c_int, c_int, vp, c_size_t, u32p, u32p, u32p, vp]),
        "zng_b200_deflate_chunks_primed": (c_int, [vp, vp, c_size_t, c_uint32, c_int, c_int, vp, c_size_t, u32p, u32p, u32p, vp]),
        "zng_b200_deflate_chunks_primed_at": (c_int, [vp, vp, c_size_t, c_uint32, c_int, c_int, c_int, vp, c_size_t, u32p, u32p, u32p, vp]),
        "zng_b200_halo_exchange": (c_int, [vp, vp, c_size_t, vp, vp]),
        "zng_b200_deflate_chunks_trace": (c_int, [vp, vp, c_size_t, c_uint32, c_int, c_int, vp, c_size_t, u32p, u32p, c_uint32, vp]),
        "zng_b200_chunk_offsets": (c_int, [vp, u32p, c_uint32, c_uint64, u64p, vp]),
        "zng_b200_gather_chunks": (c_int, [vp, vp, c_size_t, u32p, u64p, c_uint32, vp, vp]),
        "zng_b200_checksum_chunks": (c_int, [vp, vp, c_size_t, c_uint32, u32p, u32p, vp]),
        "zng_b200_crc32_fold": (c_int, [vp, u32p, c_uint32, c_uint32, c_size_t, c_uint32, u32p, vp]),
        "zng_b200_adler32_fold": (c_int, [vp, u32p, c_uint32, c_uint32, c_size_t, c_uint32, u32p, vp]),
        "zng_b200_crc32": (c_int, [vp, vp, c_size_t, c_uint32, u32p, vp]),
        "zng_b200_adler32": (c_int, [vp, vp, c_size_t, c_uint32, u32p, vp]),
        "zng_b200_deflate_host": (c_int, [vp, vp, c_size_t, c_uint32, c_int, c_int, vp, c_size_t, POINTER(c_size_t), POINTER(c_uint32), POINTER(c_uint32)]),
        "zng_b200_deflate_host_primed": (c_int, [vp, vp, vp, c_size_t, c_int, vp, c_size_t, POINTER(c_size_t), POINTER(c_uint32), POINTER(c_uint32)]),
        "zng_b200_deflate_host_primed_level": (c_int, [vp, vp, vp, c_size_t, c_int, c_int, vp, c_size_t, POINTER(c_size_t), POINTER(c_uint32), POINTER(c_uint32)]),
        "zng_b200_inflate_stream_open": (c_int, [vp, c_int, POINTER(c_void_p)]),
        "zng_b200_inflate_stream_feed": (c_int, [vp, vp, c_size_t, vp, c_size_t, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_int32), POINTER(c_uint32), POINTER(c_uint32)]),
        "zng_b200_inflate_stream_close": (None, [vp]),
        "zng_b200_crc32_host": (c_int, [vp, vp, c_size_t, c_uint32, POINTER(c_uint32)]),
        "zng_b200_adler32_host": (c_int, [vp, vp, c_size_t, c_uint32, POINTER(c_uint32)]),
        "zng_b200_comm_unique_id": (c_int, [vp, c_size_t]),
        "zng_b200_comm_create": (c_int, [vp, c_int, c_int, vp, POINTER(c_void_p)]),
        "zng_b200_comm_adopt": (c_int, [vp, vp, POINTER(c_void_p)]),
        "zng_b200_comm_destroy": (None, [vp]),
        "zng_b200_comm_size": (c_int, [vp]),
        "zng_b200_comm_rank": (c_int, [vp]),
        "zng_b200_comm_error": (c_char_p, [vp]),
        "zng_b200_stream_index_multi": (c_int, [vp, u32p, u32p, c_uint32, c_uint32, c_size_t, c_uint64, u64p, POINTER(c_uint64), POINTER(c_uint32), POINTER(c_uint64), vp]),
        "zng_b200_gzip_multi": (c_int, [vp, vp, c_size_t, c_int, vp, c_size_t, u32p, u32p, u64p, vp, c_size_t, POINTER(c_uint64), POINTER(c_uint64),
                                        POINTER(c_uint64), vp, vp, vp]),
        "zng_b200_inflate_members": (c_int, [vp, vp, u64p, c_uint32, c_int, vp, u64p, u32p, u32p, vp, u32p, u32p, vp]),
        "zng_b200_inflate_members_host": (c_int, [vp, vp, u64p, c_uint32, c_int, vp, u64p, u32p, u32p, vp, u32p, u32p]),
        "zng_b200_inflate_msg": (c_char_p, [c_uint32]),
        "zng_b200_inflate_stream_host": (c_int, [vp, vp, c_size_t, c_int, vp, c_size_t, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_uint32), POINTER(c_int32), POINTER(c_uint32)]),
        "zng_b200_op_compare256": (c_int, [vp, vp, vp, c_size_t, c_uint32, u32p, vp]),
        "zng_b200_op_longest_match": (c_int, [vp, vp, c_uint32, vp, u32p, u32p, c_uint32, u32p, u32p, vp]),
        "zng_b200_op_insert_string": (c_int, [vp, vp, vp, vp, c_uint32, c_uint32, vp]),
        "zng_b200_op_chunkmemset": (c_int, [vp, vp, c_uint32, c_uint32, c_uint32, vp]),
        "zng_b200_functable_get": (c_void_p, []),
        "zng_b200_op_longest_match_level": (c_int, [vp, vp, c_uint32, vp, u32p, u32p, c_uint32, c_int, u32p, u32p, vp]),
        "zng_b200_op_quick_insert_string": (c_int, [vp, vp, vp, vp, c_uint32, u32p, vp]),
        "zng_b200_op_slide_hash": (c_int, [vp, vp, vp, c_uint32, vp]),
        "zng_b200_crc32_copy_host": (c_int, [vp, vp, vp, c_size_t, c_uint32, POINTER(c_uint32)]),
        "zng_b200_adler32_copy_host": (c_int, [vp, vp, vp, c_size_t, c_uint32, POINTER(c_uint32)]),
        "zng_b200_update_hash": (c_uint32, [c_uint32, c_uint32]),
        "zng_b200_insert_string": (None, [vp, c_uint32, c_uint32]),
        "zng_b200_quick_insert_string": (ctypes.c_uint16, [vp, c_uint32]),
        # the C11 host library: zlib-ng's own API (include/zlib-ng.h)
        "zlibng_version": (c_char_p, []),
        "zng_deflateInit2": (c_int32, [vp, c_int32, c_int32, c_int32, c_int32, c_int32]),
        "zng_deflate": (c_int32, [vp, c_int32]),
        "zng_deflateReset": (c_int32, [vp]),
        "zng_deflateSetDictionary": (c_int32, [vp, vp, c_uint32]),
        "zng_deflateEnd": (c_int32, [vp]),
        "zng_deflateBound": (ctypes.c_ulong, [vp, ctypes.c_ulong]),
        "zng_inflateInit2": (c_int32, [vp, c_int32]),
        "zng_inflate": (c_int32, [vp, c_int32]),
        "zng_inflateReset": (c_int32, [vp]),
        "zng_inflateEnd": (c_int32, [vp]),
        "zng_compress2": (c_int32, [vp, POINTER(c_size_t), vp, c_size_t, c_int32]),
        "zng_compressBound": (c_size_t, [c_size_t]),
        "zng_uncompress": (c_int32, [vp, POINTER(c_size_t), vp, c_size_t]),
        "zng_uncompress2": (c_int32, [vp, POINTER(c_size_t), vp, POINTER(c_size_t)]),
        "zng_crc32": (c_uint32, [c_uint32, vp, c_uint32]),
        "zng_crc32_z": (c_uint32, [c_uint32, vp, c_size_t]),
        "zng_adler32": (c_uint32, [c_uint32, vp, c_uint32]),
        "zng_adler32_z": (c_uint32, [c_uint32, vp, c_size_t]),
        "zng_crc32_combine": (c_uint32, [c_uint32, c_uint32, ctypes.c_int64]),
        "zng_crc32_combine_gen": (c_uint32, [ctypes.c_int64]),
        "zng_crc32_combine_op": (c_uint32, [c_uint32, c_uint32, c_uint32]),
        "zng_adler32_combine": (c_uint32, [c_uint32, c_uint32, ctypes.c_int64]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)          # AttributeError = a declared symbol is not exported: fail loudly
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


class Crc32Fold(ctypes.Structure):
    """struct zng_b200_crc32_fold (crc32.h:8-14 layout)."""
    _fields_ = [("fold", ctypes.c_uint8 * 64), ("value", c_uint32)]


class MatchState(ctypes.Structure):
    """struct zng_b200_match_state: the deflate_state fields longest_match / insert_string / slide_hash read and write (host memory)."""
    _fields_ = [("window", c_void_p), ("window_len", c_uint32), ("head", c_void_p), ("prev", c_void_p), ("strstart", c_uint32),
                ("lookahead", c_uint32), ("match_start", c_uint32), ("level", c_int32)]


class Functable(ctypes.Structure):
    """struct zng_b200_functable (include/zng_b200.h): the 15 slots of the reference's functable_s, same order."""
    _fields_ = [("force_init", ctypes.CFUNCTYPE(None)),
                ("adler32", ctypes.CFUNCTYPE(c_uint32, c_uint32, c_void_p, c_size_t)),
                ("adler32_fold_copy", ctypes.CFUNCTYPE(c_uint32, c_uint32, c_void_p, c_void_p, c_size_t)),
                ("chunkmemset_safe", ctypes.CFUNCTYPE(c_void_p, c_void_p, c_void_p, ctypes.c_uint, ctypes.c_uint)),
                ("chunksize", ctypes.CFUNCTYPE(c_uint32)),
                ("compare256", ctypes.CFUNCTYPE(c_uint32, c_void_p, c_void_p)),
                ("crc32", ctypes.CFUNCTYPE(c_uint32, c_uint32, c_void_p, c_size_t)),
                ("crc32_fold", ctypes.CFUNCTYPE(None, POINTER(Crc32Fold), c_void_p, c_size_t, c_uint32)),
                ("crc32_fold_copy", ctypes.CFUNCTYPE(None, POINTER(Crc32Fold), c_void_p, c_void_p, c_size_t)),
                ("crc32_fold_final", ctypes.CFUNCTYPE(c_uint32, POINTER(Crc32Fold))),
                ("crc32_fold_reset", ctypes.CFUNCTYPE(c_uint32, POINTER(Crc32Fold))),
                ("inflate_fast", ctypes.CFUNCTYPE(None, c_void_p, c_uint32)),
                ("longest_match", ctypes.CFUNCTYPE(c_uint32, POINTER(MatchState), ctypes.c_uint16)),
                ("longest_match_slow", ctypes.CFUNCTYPE(c_uint32, POINTER(MatchState), ctypes.c_uint16)),
                ("slide_hash", ctypes.CFUNCTYPE(None, POINTER(MatchState)))]


def functable() -> "Functable":
    return ctypes.cast(lib().zng_b200_functable_get(), POINTER(Functable)).contents


class ZngStream(ctypes.Structure):
    """zng_stream (include/zlib-ng.h; same layout as the reference's, zlib-ng.h.in:99-119)."""
    _fields_ = [("next_in", c_void_p), ("avail_in", c_uint32), ("total_in", c_size_t), ("next_out", c_void_p),
                ("avail_out", c_uint32), ("total_out", c_size_t), ("msg", c_char_p), ("state", c_void_p),
                ("zalloc", c_void_p), ("zfree", c_void_p), ("opaque", c_void_p), ("data_type", c_int),
                ("adler", c_uint32), ("reserved", ctypes.c_ulong)]


def inflate_msg(detail: int):
    m = lib().zng_b200_inflate_msg(detail & 0xff)
    return m.decode() if m else None


def deflate_bound(chunk_len: int) -> int:
    return int(lib().zng_b200_deflate_bound(chunk_len))


def _ptr(t) -> int:
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return 0
    if hasattr(t, "data_ptr"):
        return int(t.data_ptr())
    return int(t.ctypes.data)


class Context:
    """One zng_b200_ctx (one per host thread / device), with tensor-friendly wrappers."""

    def __init__(self, device: int = -1):
        self._h = c_void_p()
        r = lib().zng_b200_ctx_create(byref(self._h), device)
        if r != 0:
            raise ZngB200Error(r, "zng_b200_ctx_create failed (no sm_100 device? there is no CPU fallback)")
        self.device = int(lib().zng_b200_ctx_device(self._h))
        self.sm_count = int(lib().zng_b200_ctx_sm_count(self._h))

    def close(self):
        if self._h:
            lib().zng_b200_ctx_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, r: int):
        if r != 0:
            raise ZngB200Error(r, lib().zng_b200_last_error(self._h).decode())

    @staticmethod
    def _stream():
        import torch
        return int(torch.cuda.current_stream().cuda_stream)

    # ---- device-resident entry points (torch uint8 CUDA tensors) -------------------------
    def alloc_chunk_outputs(self, n: int, chunk: int = CHUNK_MAX, adler: bool = True):
        import torch
        dev = torch.device("cuda", self.device)
        nchunks = (n + chunk - 1) // chunk
        stride = deflate_bound(chunk)
        slots = torch.empty(max(nchunks, 1) * stride, dtype=torch.uint8, device=dev)
        sizes = torch.zeros(max(nchunks, 1), dtype=torch.int32, device=dev)
        crcs = torch.zeros(max(nchunks, 1), dtype=torch.int32, device=dev)
        adlers = torch.zeros(max(nchunks, 1), dtype=torch.int32, device=dev) if adler else None
        return slots, stride, sizes, crcs, adlers

    def deflate_chunks(self, d_in, n: int, chunk: int, level: int, flush: int, slots, stride: int, sizes, crcs=None, adlers=None):
        self._check(lib().zng_b200_deflate_chunks(self._h, _ptr(d_in), n, chunk, level, flush, _ptr(slots), stride,
                                                  _ptr(sizes), _ptr(crcs), _ptr(adlers), self._stream()))

    def deflate_chunks_primed(self, d_in, n: int, chunk: int, level: int, flush: int, slots, stride: int, sizes, crcs=None, adlers=None):
        """pigz's dependent-chunk mode: every chunk after the first is primed with the 32768 bytes in front of it."""
        self._check(lib().zng_b200_deflate_chunks_primed(self._h, _ptr(d_in), n, chunk, level, flush, _ptr(slots), stride,
                                                         _ptr(sizes), _ptr(crcs), _ptr(adlers), self._stream()))

    def deflate_chunks_primed_at(self, d_in, n: int, chunk: int, level: int, flush: int, have_halo: bool, slots, stride: int, sizes, crcs=None, adlers=None):
        self._check(lib().zng_b200_deflate_chunks_primed_at(self._h, _ptr(d_in), n, chunk, level, flush, 1 if have_halo else 0, _ptr(slots), stride,
                                                            _ptr(sizes), _ptr(crcs), _ptr(adlers), self._stream()))

    def deflate_chunks_trace(self, d_in, n: int, chunk: int, level: int, flush: int, slots, stride: int, sizes, tokens, tok_stride: int):
        self._check(lib().zng_b200_deflate_chunks_trace(self._h, _ptr(d_in), n, chunk, level, flush, _ptr(slots), stride,
                                                        _ptr(sizes), _ptr(tokens), tok_stride, self._stream()))

    def chunk_offsets(self, sizes, nchunks: int, base: int, offsets):
        self._check(lib().zng_b200_chunk_offsets(self._h, _ptr(sizes), nchunks, base, _ptr(offsets), self._stream()))

    def gather_chunks(self, slots, stride: int, sizes, offsets, nchunks: int, dst):
        self._check(lib().zng_b200_gather_chunks(self._h, _ptr(slots), stride, _ptr(sizes), _ptr(offsets), nchunks, _ptr(dst), self._stream()))

    def checksum_chunks(self, d_in, n: int, tile: int, crcs=None, adlers=None):
        self._check(lib().zng_b200_checksum_chunks(self._h, _ptr(d_in), n, tile, _ptr(crcs), _ptr(adlers), self._stream()))

    def crc32_fold(self, crcs, ntiles: int, tile: int, n: int, init: int, result):
        self._check(lib().zng_b200_crc32_fold(self._h, _ptr(crcs), ntiles, tile, n, init, _ptr(result), self._stream()))

    def adler32_fold(self, adlers, ntiles: int, tile: int, n: int, init: int, result):
        self._check(lib().zng_b200_adler32_fold(self._h, _ptr(adlers), ntiles, tile, n, init, _ptr(result), self._stream()))

    def crc32(self, d_buf, n: int, init: int, result):
        self._check(lib().zng_b200_crc32(self._h, _ptr(d_buf), n, init, _ptr(result), self._stream()))

    def adler32(self, d_buf, n: int, init: int, result):
        self._check(lib().zng_b200_adler32(self._h, _ptr(d_buf), n, init, _ptr(result), self._stream()))

    def inflate_members(self, d_in, in_off, n_members: int, window_bits: int, d_out, out_off, sizes, checks, status, in_used=None, detail=None):
        self._check(lib().zng_b200_inflate_members(self._h, _ptr(d_in), _ptr(in_off), n_members, window_bits, _ptr(d_out), _ptr(out_off),
                                                   _ptr(sizes), _ptr(checks), _ptr(status), _ptr(in_used), _ptr(detail), self._stream()))

    def inflate_members_host(self, h_in, in_off, window_bits: int, h_out, out_off):
        """numpy in/out; returns (sizes, checks, status, in_used, detail) as numpy arrays."""
        import numpy as np
        n = len(in_off) - 1
        io = np.ascontiguousarray(in_off, dtype=np.uint64)
        oo = np.ascontiguousarray(out_off, dtype=np.uint64)
        sizes = np.zeros(max(n, 1), dtype=np.uint32); checks = np.zeros(max(n, 1), dtype=np.uint32)
        status = np.zeros(max(n, 1), dtype=np.int32); used = np.zeros(max(n, 1), dtype=np.uint32); detail = np.zeros(max(n, 1), dtype=np.uint32)
        self._check(lib().zng_b200_inflate_members_host(self._h, _ptr(h_in), _ptr(io), n, window_bits, _ptr(h_out), _ptr(oo),
                                                        _ptr(sizes), _ptr(checks), _ptr(status), _ptr(used), _ptr(detail)))
        return sizes[:n], checks[:n], status[:n], used[:n], detail[:n]

    def inflate_stream_host(self, h_in, n: int, window_bits: int, h_out, cap: int):
        """One stream (parallel over its flush markers when it has them).  Returns (status, out_len, in_used, check, detail)."""
        ol, iu, chk, st, det = c_size_t(0), c_size_t(0), c_uint32(0), c_int32(0), c_uint32(0)
        self._check(lib().zng_b200_inflate_stream_host(self._h, _ptr(h_in), n, window_bits, _ptr(h_out), cap, byref(ol), byref(iu), byref(chk), byref(st), byref(det)))
        return int(st.value), int(ol.value), int(iu.value), int(chk.value), int(det.value)

    # ---- host-buffer entry points (numpy arrays / pinned tensors) ------------------------
    def deflate_host(self, h_in, n: int, chunk: int, level: int, final: bool, h_out, out_cap: int):
        out_len = c_size_t(0)
        crc = c_uint32(0)
        adler = c_uint32(0)
        self._check(lib().zng_b200_deflate_host(self._h, _ptr(h_in), n, chunk, level, 1 if final else 0, _ptr(h_out), out_cap,
                                                byref(out_len), byref(crc), byref(adler)))
        return int(out_len.value), int(crc.value), int(adler.value)

    def crc32_host(self, h_buf, n: int, init: int = 0) -> int:
        res = c_uint32(0)
        self._check(lib().zng_b200_crc32_host(self._h, _ptr(h_buf), n, init, byref(res)))
        return int(res.value)

    def adler32_host(self, h_buf, n: int, init: int = 1) -> int:
        res = c_uint32(0)
        self._check(lib().zng_b200_adler32_host(self._h, _ptr(h_buf), n, init, byref(res)))
        return int(res.value)


class Comm:
    """One zng_b200_comm (csrc/multi.cu): the NCCL communicator of the multi-GPU stream assembly.  `Comm.from_torch(ctx)`
    builds it for the ranks of the default torch.distributed process group (the 128-byte NCCL id travels by broadcast)."""

    def __init__(self, ctx: "Context", nranks: int, rank: int, uid: bytes | None):
        self.ctx = ctx
        self._h = c_void_p()
        buf = ctypes.create_string_buffer(uid, 128) if uid is not None else None
        r = lib().zng_b200_comm_create(ctx._h, nranks, rank, buf, byref(self._h))
        if r != 0:
            raise ZngB200Error(r, "zng_b200_comm_create failed (NCCL not loadable, or bad arguments)")
        self.nranks, self.rank = nranks, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        r = lib().zng_b200_comm_unique_id(buf, 128)
        if r != 0:
            raise ZngB200Error(r, "zng_b200_comm_unique_id failed (NCCL not loadable)")
        return buf.raw

    @classmethod
    def from_torch(cls, ctx: "Context"):
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return cls(ctx, 1, 0, None)
        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{ctx.device}")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return cls(ctx, world, rank, bytes(t.cpu().numpy().tobytes()))

    def close(self):
        if self._h:
            lib().zng_b200_comm_destroy(self._h)
            self._h = c_void_p()

    def _check(self, r: int):
        if r != 0:
            raise ZngB200Error(r, lib().zng_b200_comm_error(self._h).decode() or lib().zng_b200_last_error(self.ctx._h).decode())

    def halo_exchange(self, d_in, n_local: int, d_halo):
        """Collective: this rank's last 32 KiB to the next rank, the previous rank's into d_halo (directly in front of d_in)."""
        self._check(lib().zng_b200_halo_exchange(self._h, _ptr(d_in), n_local, _ptr(d_halo), Context._stream()))

    def stream_index(self, sizes, crcs, nchunks_local: int, chunk: int, n_local: int, base: int, offsets_local):
        """Collective.  Returns (stream_end, crc32, total_in); offsets_local (int64 device tensor, nchunks_local + 1) is filled."""
        end, crc, tin = c_uint64(0), c_uint32(0), c_uint64(0)
        self._check(lib().zng_b200_stream_index_multi(self._h, _ptr(sizes), _ptr(crcs), nchunks_local, chunk, n_local, base,
                                                      _ptr(offsets_local), byref(end), byref(crc), byref(tin), Context._stream()))
        return int(end.value), int(crc.value), int(tin.value)

    def gzip_multi(self, d_in, n_local: int, level: int, slots, stride: int, sizes, crcs, offsets_local, packed, packed_cap: int):
        """Collective: compress + allgather + scan + fold + pack.  Returns (my_offset, my_bytes, file_bytes, header, trailer)."""
        off, nb, fb = c_uint64(0), c_uint64(0), c_uint64(0)
        hdr = ctypes.create_string_buffer(10); trl = ctypes.create_string_buffer(10)
        self._check(lib().zng_b200_gzip_multi(self._h, _ptr(d_in), n_local, level, _ptr(slots), stride, _ptr(sizes), _ptr(crcs), _ptr(offsets_local),
                                              _ptr(packed), packed_cap, byref(off), byref(nb), byref(fb), hdr, trl, Context._stream()))
        return int(off.value), int(nb.value), int(fb.value), hdr.raw, trl.raw
