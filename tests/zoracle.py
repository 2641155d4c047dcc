"""ctypes bindings of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

`port_*`  : oracle/libzng_oracle.so, our plain-C restatement of the reference algorithms
            (oracle/zo_*.c), always available (built by __graft_entry__.build()).
`ref_*`   : oracle/_ref/libzng_ref.so, the UNMODIFIED reference compiled from /root/reference by
            oracle/Makefile, present when it was built in the authoring container (it travels to
            the GPU box as a built .so).  `have_ref()` says whether it is there.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (zlib-ng_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_PATH = os.path.join(ROOT, "oracle", "libzng_oracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libzng_ref.so")

_port = None
_ref = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_PATH):
            raise ImportError(f"{PORT_PATH} missing: run `make -C oracle port`")
        L = ctypes.CDLL(PORT_PATH)
        L.zo_crc32.restype = c_uint32; L.zo_crc32.argtypes = [c_uint32, c_void_p, c_size_t]
        L.zo_adler32.restype = c_uint32; L.zo_adler32.argtypes = [c_uint32, c_void_p, c_size_t]
        L.zo_crc32_combine.restype = c_uint32; L.zo_crc32_combine.argtypes = [c_uint32, c_uint32, c_int64]
        L.zo_crc32_combine_gen.restype = c_uint32; L.zo_crc32_combine_gen.argtypes = [c_int64]
        L.zo_crc32_combine_op.restype = c_uint32; L.zo_crc32_combine_op.argtypes = [c_uint32, c_uint32, c_uint32]
        L.zo_adler32_combine.restype = c_uint32; L.zo_adler32_combine.argtypes = [c_uint32, c_uint32, c_int64]
        L.zo_compare256.restype = c_uint32; L.zo_compare256.argtypes = [c_void_p, c_void_p]
        L.zo_frame_gzip_members.restype = c_size_t
        L.zo_frame_gzip_members.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_uint32, c_uint32, c_size_t, c_int, c_void_p, c_void_p]
        L.zo_compare_chunks.restype = c_size_t
        L.zo_compare_chunks.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t)]
        L.zo_deflate_bound.restype = c_size_t; L.zo_deflate_bound.argtypes = [c_size_t]
        L.zo_deflate_chunk.restype = c_size_t; L.zo_deflate_chunk.argtypes = [c_void_p, c_uint32, c_int, c_int, c_void_p, c_size_t]
        L.zo_deflate_chunks.restype = c_int
        L.zo_deflate_chunks.argtypes = [c_void_p, c_size_t, c_uint32, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_int]
        L.zo_deflate_chunks_primed.restype = c_int
        L.zo_deflate_chunks_primed.argtypes = L.zo_deflate_chunks.argtypes
        L.zo_deflate_chunks_fresh_window.restype = c_int
        L.zo_deflate_chunks_fresh_window.argtypes = L.zo_deflate_chunks.argtypes
        L.zo_deflate_tokens.restype = c_size_t; L.zo_deflate_tokens.argtypes = [c_void_p, c_uint32, c_int, c_void_p, c_size_t]
        L.zo_deflate_tokens_primed.restype = c_size_t; L.zo_deflate_tokens_primed.argtypes = [c_void_p, c_uint32, c_int, c_void_p, c_size_t]
        L.zo_longest_match_l2.restype = c_uint32
        L.zo_longest_match_l2.argtypes = [c_void_p, c_uint32, c_uint32, c_void_p, c_uint32, c_uint32, POINTER(c_uint32)]
        L.zo_insert_string.restype = None
        L.zo_insert_string.argtypes = [c_void_p, c_uint32, c_void_p, c_void_p, c_uint32, c_uint32]
        if hasattr(L, "zo_inflate"):
            L.zo_inflate.restype = c_int
            L.zo_inflate.argtypes = [c_void_p, c_size_t, c_int, c_void_p, c_size_t, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_uint32), POINTER(c_char_p)]
            L.zo_inflate_members.restype = c_int
            L.zo_inflate_members.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
        _port = L
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise ImportError(f"{REF_PATH} missing (built from /root/reference by `make -C oracle ref`)")
        L = ctypes.CDLL(REF_PATH)
        L.refdrv_deflate_chunks.restype = c_int
        L.refdrv_deflate_chunks.argtypes = [c_void_p, c_size_t, c_uint32, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_int]
        if hasattr(L, "refdrv_deflate_chunks_primed"):
            L.refdrv_deflate_chunks_primed.restype = c_int
            L.refdrv_deflate_chunks_primed.argtypes = L.refdrv_deflate_chunks.argtypes
        L.refdrv_checksum_chunks.restype = c_int
        L.refdrv_checksum_chunks.argtypes = [c_void_p, c_size_t, c_uint32, c_void_p, c_void_p, c_int]
        L.refdrv_checksum_flat.restype = c_int
        L.refdrv_checksum_flat.argtypes = [c_void_p, c_size_t, c_uint32, c_int, POINTER(c_uint32), POINTER(c_uint32)]
        L.refdrv_inflate_members.restype = c_int
        L.refdrv_inflate_members.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
        L.refdrv_gzip_members.restype = c_int
        L.refdrv_gzip_members.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_void_p, c_int]
        L.refdrv_inflate_stream.restype = c_int
        L.refdrv_inflate_stream.argtypes = [c_void_p, c_size_t, c_int, c_void_p, c_size_t, POINTER(c_uint64), POINTER(c_uint32)]
        if hasattr(L, "refops_longest_match"):          # oracle/ref_ops.c: the reference's operators one at a time
            L.refops_longest_match.restype = c_uint32
            L.refops_longest_match.argtypes = [c_void_p, c_uint32, c_void_p, c_uint32, c_uint32, c_int, POINTER(c_uint32)]
            L.refops_insert_string.restype = c_uint32
            L.refops_insert_string.argtypes = [c_void_p, c_uint32, c_void_p, c_void_p, c_uint32, c_uint32]
            L.refops_slide_hash.restype = c_int; L.refops_slide_hash.argtypes = [c_void_p, c_void_p]
            L.refops_update_hash.restype = c_uint32; L.refops_update_hash.argtypes = [c_uint32, c_uint32]
            L.refops_compare256.restype = c_uint32; L.refops_compare256.argtypes = [c_void_p, c_void_p]
            L.refops_chunksize.restype = c_uint32; L.refops_chunksize.argtypes = []
            L.refops_chunkmemset_safe.restype = c_uint32; L.refops_chunkmemset_safe.argtypes = [c_void_p, c_uint32, c_uint32, c_uint32, c_uint32]
            L.refops_crc32_fold.restype = c_uint32; L.refops_crc32_fold.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p]
            L.refops_adler32_fold_copy.restype = c_uint32; L.refops_adler32_fold_copy.argtypes = [c_uint32, c_void_p, c_void_p, c_size_t]
        L.refdrv_now.restype = c_double
        L.refdrv_inflate_oneshot.restype = c_int
        L.refdrv_inflate_oneshot.argtypes = [c_void_p, c_size_t, c_int, c_void_p, c_size_t, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_uint32), c_char_p]
        L.refdrv_deflate_stream.restype = c_int
        L.refdrv_deflate_stream.argtypes = [c_void_p, c_size_t, c_uint32, c_int, c_int, c_void_p, c_size_t, POINTER(c_size_t)]
        for name, res, args in (
            ("zng_crc32", c_uint32, [c_uint32, c_void_p, c_uint32]),
            ("zng_crc32_z", c_uint32, [c_uint32, c_void_p, c_size_t]),
            ("zng_adler32", c_uint32, [c_uint32, c_void_p, c_uint32]),
            ("zng_adler32_z", c_uint32, [c_uint32, c_void_p, c_size_t]),
            ("zng_crc32_combine", c_uint32, [c_uint32, c_uint32, c_int64]),
            ("zng_crc32_combine_gen", c_uint32, [c_int64]),
            ("zng_crc32_combine_op", c_uint32, [c_uint32, c_uint32, c_uint32]),
            ("zng_adler32_combine", c_uint32, [c_uint32, c_uint32, c_int64]),
        ):
            f = getattr(L, name); f.restype = res; f.argtypes = args
        _ref = L
    return _ref


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


_ONE = np.zeros(16, dtype=np.uint8)


def _ptr(a: np.ndarray):
    # an empty array still gets a valid (non-NULL) pointer: NULL means "no buffer" to zng_crc32
    return c_void_p(a.ctypes.data) if a.size else c_void_p(_ONE.ctypes.data)


def _deflate_chunks(fn, data, chunk, level, flush, stride, nthreads):
    data = _u8(data)
    n = data.size
    nch = (n + chunk - 1) // chunk
    out = np.zeros((max(nch, 1), stride), dtype=np.uint8)
    sizes = np.zeros(max(nch, 1), dtype=np.uint32)
    crcs = np.zeros(max(nch, 1), dtype=np.uint32)
    adlers = np.zeros(max(nch, 1), dtype=np.uint32)
    r = fn(_ptr(data), n, chunk, level, flush, out.ctypes.data, stride, sizes.ctypes.data, crcs.ctypes.data, adlers.ctypes.data, nthreads)
    if r != 0:
        raise RuntimeError(f"oracle deflate_chunks failed: {r}")
    return out[:nch], sizes[:nch], crcs[:nch], adlers[:nch]


def gzip_members(data, member=4096, level=1, nthreads=None):
    """Every `member`-byte slice of data as one gzip member (body = the level-`level` Z_FINISH stream of the slice written by the
    strongest checker available, framing as minigzip writes it), packed.  Returns (members u8 (+16 bytes of slack), off u64[n+1])."""
    data = _u8(data)
    _, out, sizes, crcs, _ = best_deflate_chunks(data, member, level, 4, None, nthreads)
    nm = sizes.size
    last = data.size - (nm - 1) * member
    off = np.zeros(nm + 1, dtype=np.uint64)
    xfl = 4 if level == 1 else 0
    total = port().zo_frame_gzip_members(out.ctypes.data, out.shape[1], sizes.ctypes.data, crcs.ctypes.data, member, last, nm, xfl, None, None)
    buf = np.zeros(total + 16, dtype=np.uint8)
    port().zo_frame_gzip_members(out.ctypes.data, out.shape[1], sizes.ctypes.data, crcs.ctypes.data, member, last, nm, xfl, buf.ctypes.data, off.ctypes.data)
    return buf, off


def compare_chunks(a, stride_a, sizes_a, b, stride_b, sizes_b):
    """Every chunk of `a` against `b`, byte for byte (stride 0 = packed back to back).  Returns (bad_count, first_bad)."""
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1); b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    sa = np.ascontiguousarray(sizes_a, dtype=np.uint32); sb = np.ascontiguousarray(sizes_b, dtype=np.uint32)
    assert sa.size == sb.size
    first = c_size_t(0)
    bad = port().zo_compare_chunks(_ptr(a), stride_a, sa.ctypes.data, _ptr(b), stride_b, sb.ctypes.data, sa.size, byref(first))
    return int(bad), int(first.value)


def best_deflate_chunks(data, chunk=65536, level=1, flush=3, stride=None, nthreads=None):
    """The strongest checker available: the unmodified reference when oracle/_ref is present, else the port.
    Returns (kind, out, sizes, crcs, adlers)."""
    if have_ref():
        return ("oracle/_ref (unmodified zlib-ng 2.2.2)",) + ref_deflate_chunks(data, chunk, level, flush, stride, nthreads)
    return ("oracle port",) + port_deflate_chunks(data, chunk, level, flush, stride, nthreads)


def port_deflate_chunks(data, chunk=65536, level=1, flush=3, stride=None, nthreads=None):
    stride = stride or int(port().zo_deflate_bound(chunk))
    return _deflate_chunks(port().zo_deflate_chunks, data, chunk, level, flush, stride, nthreads or min(os.cpu_count() or 1, 32))


def ref_deflate_chunks(data, chunk=65536, level=1, flush=3, stride=None, nthreads=None):
    stride = stride or int(port().zo_deflate_bound(chunk))
    return _deflate_chunks(ref().refdrv_deflate_chunks, data, chunk, level, flush, stride, nthreads or min(os.cpu_count() or 1, 32))


def port_deflate_chunks_primed(data, chunk=65536, level=1, flush=2, stride=None, nthreads=None):
    """pigz's dependent mode: chunk u > 0 primed with the 32768 stream bytes in front of it (fresh stream + deflateSetDictionary)."""
    stride = stride or int(port().zo_deflate_bound(chunk))
    return _deflate_chunks(port().zo_deflate_chunks_primed, data, chunk, level, flush, stride, nthreads or min(os.cpu_count() or 1, 32))


def port_deflate_chunks_fresh_window(data, chunk=65536, level=2, flush=4, stride=None, nthreads=None):
    """Levels 2-6, every chunk on a fresh stream, through the oracle's window engine (second restatement)."""
    stride = stride or int(port().zo_deflate_bound(chunk))
    return _deflate_chunks(port().zo_deflate_chunks_fresh_window, data, chunk, level, flush, stride, nthreads or min(os.cpu_count() or 1, 32))


def ref_deflate_chunks_primed(data, chunk=65536, level=1, flush=2, stride=None, nthreads=None):
    stride = stride or int(port().zo_deflate_bound(chunk))
    return _deflate_chunks(ref().refdrv_deflate_chunks_primed, data, chunk, level, flush, stride, nthreads or min(os.cpu_count() or 1, 32))


def port_tokens_primed(dict_and_chunk, level=1) -> np.ndarray:
    """LZ77 tokens of a primed chunk; the first 32768 bytes of the argument are its dictionary."""
    d = _u8(dict_and_chunk)
    assert d.size >= 32768
    n = d.size - 32768
    tok = np.zeros(n + 8, dtype=np.uint32)
    k = port().zo_deflate_tokens_primed(d.ctypes.data + 32768, n, level, tok.ctypes.data, tok.size)
    return tok[:k]


def port_tokens(chunk_bytes, level=1) -> np.ndarray:
    d = _u8(chunk_bytes)
    tok = np.zeros(d.size + 8, dtype=np.uint32)
    k = port().zo_deflate_tokens(_ptr(d), d.size, level, tok.ctypes.data, tok.size)
    return tok[:k]


def port_crc32(data, init=0) -> int:
    d = _u8(data)
    return int(port().zo_crc32(init, _ptr(d), d.size))


def port_adler32(data, init=1) -> int:
    d = _u8(data)
    return int(port().zo_adler32(init, _ptr(d), d.size))


def ref_crc32(data, init=0) -> int:
    d = _u8(data)
    return int(ref().zng_crc32_z(init, _ptr(d), d.size))


def ref_adler32(data, init=1) -> int:
    d = _u8(data)
    return int(ref().zng_adler32_z(init, _ptr(d), d.size))


def ref_inflate_stream(stream, window_bits, expect=None):
    s = _u8(stream)
    e = _u8(expect) if expect is not None else None
    total = c_uint64(0)
    crc = c_uint32(0)
    r = ref().refdrv_inflate_stream(_ptr(s), s.size, window_bits, _ptr(e) if e is not None else c_void_p(0),
                                    e.size if e is not None else 0, byref(total), byref(crc))
    return int(r), int(total.value), int(crc.value)


def _inflate(fn_kind, stream, window_bits, out_cap):
    s = _u8(stream)
    out = np.zeros(max(out_cap, 1), dtype=np.uint8)
    ol, iu, chk = c_size_t(0), c_size_t(0), c_uint32(0)
    if fn_kind == "ref":
        msg = ctypes.create_string_buffer(64)
        r = ref().refdrv_inflate_oneshot(_ptr(s), s.size, window_bits, out.ctypes.data, out_cap, byref(ol), byref(iu), byref(chk), msg)
        m = msg.value.decode() or None
    else:
        mp = c_char_p()
        r = port().zo_inflate(_ptr(s), s.size, window_bits, out.ctypes.data, out_cap, byref(ol), byref(iu), byref(chk), byref(mp))
        m = mp.value.decode() if mp.value else None
    return int(r), out[: ol.value], int(iu.value), int(chk.value), m


def port_inflate(stream, window_bits, out_cap):
    """(ret, output, in_used, check, msg) of the oracle restatement of one zng_inflate(Z_FINISH)."""
    return _inflate("port", stream, window_bits, out_cap)


def ref_inflate(stream, window_bits, out_cap):
    return _inflate("ref", stream, window_bits, out_cap)


def ref_deflate_stream(data, piece, level, window_bits):
    d = _u8(data)
    cap = d.size + d.size // 8 + (d.size // max(piece, 1) + 2) * 16 + 64
    out = np.zeros(cap, dtype=np.uint8)
    ol = c_size_t(0)
    r = ref().refdrv_deflate_stream(_ptr(d), d.size, piece, level, window_bits, out.ctypes.data, cap, byref(ol))
    if r != 1:
        raise RuntimeError(f"reference deflate stream failed: {r}")
    return out[: ol.value]
