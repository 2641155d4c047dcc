#!/usr/bin/env python3
"""Generate tests/golden/*.json -- run in the authoring container (needs /root/reference and
oracle/_ref/libzng_ref.so); the outputs are committed so that the GPU box, which has neither,
can still check parity.

  kat_crc32.json / kat_adler32.json
      The known-answer vectors of the reference's own unit tests, extracted from the C
      initialisers in test/test_crc32.cc:29-183 and test/test_adler32.cc:26-345.
  deflate_digests.json
  primed_digests.json
      Per-chunk (compressed size, crc32 of the compressed bytes, crc32 and adler32 of the input)
      produced by the UNMODIFIED reference (oracle/_ref, zlib-ng 2.2.2, gcc -O3, default build
      flags, x86-64 => OPTIMAL_CMP 64) for seeded inputs of the synthetic generator
      (tests/synth.c) and a few hand-made edge cases.  Compressed bytes are pinned by
      no test of the reference, so these digests are the pin.
  inflate_kat.json
      The hand-written bitstreams of test/infcover.c with the return code the reference expects.
"""
from __future__ import annotations

import json
import os
import re
import sys
import zlib as _pyzlib  # only for crc32-of-bytes digests of the golden file itself (not a parity oracle)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synthdata import synth  # noqa: E402


def c_unescape(lit: str) -> bytes:
    """Decode the inside of one C string literal."""
    out = bytearray()
    i = 0
    simple = {"n": 10, "t": 9, "r": 13, "0": 0, "\\": 92, '"': 34, "'": 39, "a": 7, "b": 8, "f": 12, "v": 11, "?": 63}
    while i < len(lit):
        c = lit[i]
        if c != "\\":
            out.append(ord(c)); i += 1; continue
        i += 1
        c = lit[i]
        if c == "x":
            j = i + 1
            while j < len(lit) and lit[j] in "0123456789abcdefABCDEF":
                j += 1
            out.append(int(lit[i + 1:j], 16) & 0xff); i = j
        elif c in "01234567":
            j = i
            while j < len(lit) and j < i + 3 and lit[j] in "01234567":
                j += 1
            out.append(int(lit[i:j], 8) & 0xff); i = j
        else:
            out.append(simple[c]); i += 1
    return bytes(out)


STR_RE = re.compile(r'"((?:[^"\\]|\\.)*)"')


def parse_vectors(path: str, start_marker: str, long_string: bytes | None):
    src = open(path, encoding="latin-1").read()
    body = src[src.index(start_marker):]
    body = body[: body.index("};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    vecs = []
    # each entry: {init, (const uint8_t *)EXPR, len, expect}
    ent = re.compile(r"\{\s*(0x[0-9a-fA-F]+)\s*,\s*\(const uint8_t \*\)\s*(.*?)\s*,\s*(\d+)\s*,\s*(0x[0-9a-fA-F]+)\s*\}", re.S)
    for m in ent.finditer(body):
        init, expr, ln, expect = int(m.group(1), 16), m.group(2), int(m.group(3)), int(m.group(4), 16)
        if expr.strip() == "0x0":
            data = None
        elif expr.strip() == "long_string":
            data = long_string
        else:
            data = b"".join(c_unescape(s) for s in STR_RE.findall(expr)) + b"\x00"   # literal's NUL is addressable
        vecs.append({"init": init, "data_hex": None if data is None else data[:max(ln, 0)].hex(), "len": ln, "expect": expect})
    return vecs


def parse_long_string(path: str) -> bytes:
    src = open(path, encoding="latin-1").read()
    body = src[src.index("long_string[5552]"):]
    body = body[body.index("{") + 1: body.index("};")]
    chars = re.findall(r"'((?:[^'\\]|\\.)+)'", body)
    data = b"".join(c_unescape(c) for c in chars)
    assert len(data) == 5552, len(data)
    return data


def parse_infcover(path: str):
    """inf("hex bytes", "what", step, win, len, err) calls of test/infcover.c."""
    src = open(path, encoding="latin-1").read()
    out = []
    call = re.compile(r'\binf\(\s*((?:"(?:[^"\\]|\\.)*"\s*)+),\s*"((?:[^"\\]|\\.)*)"\s*,\s*(\d+)\s*,\s*(-?\d+)\s*,\s*(\d+)\s*,\s*([A-Z_a-z0-9]+)\s*\)')
    codes = {"Z_OK": 0, "Z_STREAM_END": 1, "Z_NEED_DICT": 2, "Z_DATA_ERROR": -3, "Z_BUF_ERROR": -5, "Z_MEM_ERROR": -4, "Z_STREAM_ERROR": -2}
    for m in call.finditer(src):
        hexs = "".join(c_unescape(s).decode("latin-1") for s in STR_RE.findall(m.group(1)))
        err = m.group(6)
        out.append({"hex": hexs, "what": m.group(2), "step": int(m.group(3)), "win": int(m.group(4)),
                    "len": int(m.group(5)), "err": codes.get(err, err)})
    return out


def deflate_digests():
    import numpy as np
    from __graft_entry__ import load_oracle, load_package
    pkg = load_package()
    zo = load_oracle()
    cases = []

    def add(name, data, chunk, level, flush):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out, sizes, crcs, adlers = zo.ref_deflate_chunks(data, chunk, level, flush, nthreads=1)   # stream order: see SURVEY 0.6
        comp_crc = [int(_pyzlib.crc32(out[i, : sizes[i]].tobytes())) for i in range(len(sizes))]
        cases.append({"name": name, "chunk": chunk, "level": level, "flush": flush, "n": int(data.size),
                      "sizes": [int(x) for x in sizes], "comp_crc32": comp_crc,
                      "crc32": [int(x) for x in crcs], "adler32": [int(x) for x in adlers]})

    rng = np.random.default_rng(20261018)
    for level in (1, 2, 3, 4, 5, 6):
        add(f"synth_2MiB_l{level}", synth(32 * 65536), 65536, level, 3)
        add(f"synth_ragged_l{level}", synth(10 * 65536 + 777, seed=12345), 65536, level, 3)
        add(f"synth_finish_l{level}", synth(4 * 65536 + 4097, seed=99), 65536, level, 4)
        add(f"synth_4k_members_l{level}", synth(64 * 4096, seed=7), 4096, level, 4)
        add(f"zeros_l{level}", np.zeros(3 * 65536 + 5, dtype=np.uint8), 65536, level, 3)
        add(f"random_l{level}", rng.integers(0, 256, size=2 * 65536 + 100, dtype=np.uint8), 65536, level, 3)
        add(f"tiny_sizes_l{level}", synth(65536, seed=3)[:257 * 40], 257, level, 3)
        for n in (0, 1, 2, 3, 4, 5, 7, 8, 9, 258, 259, 260, 262, 263):
            add(f"short_{n}_l{level}", synth(65536, seed=5)[:n], 65536, level, 3)
            add(f"short_{n}_finish_l{level}", synth(65536, seed=5)[:n], 65536, level, 4)
    return cases


def primed_digests():
    """pigz's dependent-chunk mode: per chunk a fresh stream + zng_deflateSetDictionary(32768 bytes in front) + zng_deflate(flush)."""
    from __graft_entry__ import load_oracle, load_package
    pkg = load_package()
    zo = load_oracle()
    cases = []
    for seed, n, flush in ((61, 16 * 65536 + 4321, 2), (62, 8 * 65536, 3), (63, 3 * 65536 + 1, 4), (64, 2 * 65536 + 261, 2), (65, 2 * 65536 + 262, 2),
                           (66, 2 * 65536 + 263, 4), (67, 65536 + 32768, 2), (68, 65536 + 32769, 2), (69, 65536 + 65274, 4), (70, 65536 + 65275, 2),
                           (71, 65536 + 65535, 2), (72, 65536 + 3, 2), (73, 4097, 4)):
        data = synth(n, seed=seed)
        for level in (1, 2, 3, 4, 5, 6):
            if level > 1 and n > 9 * 65536:
                continue
            out, sizes, _, _ = zo.ref_deflate_chunks_primed(data, 65536, level, flush)
            cases.append({"seed": seed, "n": n, "flush": flush, "level": level, "sizes": [int(x) for x in sizes],
                          "comp_crc32": [int(_pyzlib.crc32(out[i, : sizes[i]].tobytes())) for i in range(len(sizes))]})
    return cases


def main():
    long_string = parse_long_string(f"{REF}/test/test_adler32.cc")
    crc = parse_vectors(f"{REF}/test/test_crc32.cc", "static const crc32_test tests[]", None)
    adl = parse_vectors(f"{REF}/test/test_adler32.cc", "static const adler32_test tests[]", long_string)
    assert len(crc) >= 100 and len(adl) >= 100, (len(crc), len(adl))
    json.dump({"source": "test/test_crc32.cc:29-183", "vectors": crc}, open(f"{HERE}/kat_crc32.json", "w"), indent=0)
    json.dump({"source": "test/test_adler32.cc:26-345", "vectors": adl}, open(f"{HERE}/kat_adler32.json", "w"), indent=0)
    inf = parse_infcover(f"{REF}/test/infcover.c")
    json.dump({"source": "test/infcover.c inf() calls", "vectors": inf}, open(f"{HERE}/inflate_kat.json", "w"), indent=0)
    dd = deflate_digests()
    json.dump({"source": "oracle/_ref (unmodified zlib-ng 2.2.2) via refdrv_deflate_chunks; inputs from tests/synth.c",
               "cases": dd}, open(f"{HERE}/deflate_digests.json", "w"))
    pd = primed_digests()
    json.dump({"source": "oracle/_ref (unmodified zlib-ng 2.2.2) via refdrv_deflate_chunks_primed, levels 1-6; inputs synth(n, seed)",
               "cases": pd}, open(f"{HERE}/primed_digests.json", "w"))
    print(f"primed digest cases {len(pd)}")
    print(f"crc32 KATs {len(crc)}, adler32 KATs {len(adl)}, infcover vectors {len(inf)}, deflate digest cases {len(dd)}")


if __name__ == "__main__":
    main()
