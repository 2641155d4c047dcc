"""The host-buffer entry point (what zng_deflate of the host library calls): pipelined H2D ->
K1 -> gather -> D2H, compared with the oracle's concatenated chunk stream."""
import zlib as pyzlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def oracle_stream(zo, data, chunk, level, final):
    n = data.size
    nch = (n + chunk - 1) // chunk
    parts = []
    if final:
        body = (nch - 1) * chunk if nch else 0
        if body:
            out, sizes, _, _ = zo.port_deflate_chunks(data[:body], chunk, level, 3)
            parts += [out[i, : sizes[i]].tobytes() for i in range(len(sizes))]
        if n - body:
            out, sizes, _, _ = zo.port_deflate_chunks(data[body:], chunk, level, 4)
            parts.append(out[0, : sizes[0]].tobytes())
        else:
            parts.append(b"\x03\x00")
    else:
        out, sizes, _, _ = zo.port_deflate_chunks(data, chunk, level, 3)
        parts += [out[i, : sizes[i]].tobytes() for i in range(len(sizes))]
    return b"".join(parts)


@pytest.mark.parametrize("n,final", [(0, True), (0, False), (1, True), (65536, True), (65536, False), (5 * 65536 + 99, True),
                                     (5 * 65536 + 99, False), ((40 << 20) + 12345, True), ((100 << 20), False)])
def test_deflate_host_matches_oracle_stream(pkg, ctx, zo, n, final):
    data = pkg.synth(n, seed=n % 1000 + 1)
    cap = pkg.deflate_bound(65536) * ((n + 65535) // 65536 + 1)
    out = np.empty(cap, dtype=np.uint8)
    out_len, crc, adler = ctx.deflate_host(data, n, 65536, 1, final, out, cap)
    exp = oracle_stream(zo, data, 65536, 1, final)
    assert out_len == len(exp)
    assert out[:out_len].tobytes() == exp
    assert crc == pyzlib.crc32(data.tobytes()) and adler == pyzlib.adler32(data.tobytes())
    if final:
        assert pyzlib.decompress(out[:out_len].tobytes(), wbits=-15) == data.tobytes()


def test_deflate_host_output_too_small(pkg, ctx):
    data = pkg.synth(4 * 65536, seed=3)
    out = np.empty(1000, dtype=np.uint8)
    with pytest.raises(pkg.ZngB200Error) as ei:
        ctx.deflate_host(data, data.size, 65536, 1, True, out, 1000)
    assert ei.value.code == pkg.Z_BUF_ERROR
