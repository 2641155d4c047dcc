#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: read a gzip file back with the UNMODIFIED reference's zng_inflate (oracle/_ref) and print the return
code, the output length and its CRC-32 -- how the 64 GiB / 8-GPU stream of zlib-ng_b200/tools/pigz_multi_gpu.py was checked
(profiles/r1_config5_64GiB_8gpu.txt).        python tests/verify_gzip_with_reference.py FILE [expected_len expected_crc32_hex]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    from __graft_entry__ import load_oracle
    zo = load_oracle()
    if not zo.have_ref():
        sys.exit("oracle/_ref is not built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)")
    t0 = time.perf_counter()
    data = np.fromfile(sys.argv[1], dtype=np.uint8)
    code, out_len, crc = zo.ref_inflate_stream(data, 31)
    print(f"reference zng_inflate: code {code}, {out_len} bytes, crc32 {crc & 0xffffffff:08x} ({time.perf_counter() - t0:.1f} s)")
    ok = code == 1
    if len(sys.argv) >= 4:
        ok = ok and out_len == int(sys.argv[2]) and (crc & 0xffffffff) == int(sys.argv[3], 16)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
