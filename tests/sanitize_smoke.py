#!/usr/bin/env python3
"""The smoke set for compute-sanitizer (memcheck / racecheck / synccheck, one tool per run):

    compute-sanitizer --tool memcheck python tests/sanitize_smoke.py

Small inputs through every kernel family -- K1 (both parsers), K1 primed, K2 levels 2 and 6, K2w (primed levels 2 and 5), K3, K4,
the operator kernels and the streamed host path -- each checked against the oracle so that a sanitizer run is also a parity run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from __graft_entry__ import load_package, load_oracle
from synthdata import synth

pkg = load_package(); zo = load_oracle()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)


def chunks(ctx, data, level, primed=False):
    n = data.size
    d_in = torch.from_numpy(data).to(dev)
    slots, stride, sizes, crcs, adlers = ctx.alloc_chunk_outputs(n, 65536, adler=True)
    (ctx.deflate_chunks_primed if primed else ctx.deflate_chunks)(d_in, n, 65536, level, 2 if primed else 3, slots, stride, sizes, crcs, adlers)
    torch.cuda.synchronize()
    nch = (n + 65535) // 65536
    exp, es, ec, ea = (zo.port_deflate_chunks_primed if primed else zo.port_deflate_chunks)(data, 65536, level, 2 if primed else 3, stride)
    bad, first = zo.compare_chunks(slots.cpu().numpy(), stride, sizes.cpu().numpy().view(np.uint32)[:nch], exp, stride, es)
    assert bad == 0, (level, primed, bad, first)
    assert np.array_equal(crcs.cpu().numpy().view(np.uint32)[:nch], ec)


data = synth(6 * 65536 + 12345, seed=17)
for env in (("warp", "cta") if which in ("all", "k1") else ()):
    os.environ["ZNG_B200_K1"] = env
    ctx = pkg.Context(0)
    chunks(ctx, data, 1)
    ctx.close()
    print(f"K1 ({env}) ok", flush=True)
os.environ["ZNG_B200_K1"] = "warp"
ctx = pkg.Context(0)
if which in ("all", "k2"):
    for level in (2, 6):
        chunks(ctx, data, level)
        print(f"K2 level {level} ok", flush=True)
if which in ("all", "primed"):
    for level in (1, 2, 5):
        chunks(ctx, data[: 3 * 65536 + 777], level, primed=True)
        print(f"primed level {level} ok", flush=True)
if which in ("all", "k34"):
    d_in = torch.from_numpy(data).to(dev)
    res = torch.zeros(4, dtype=torch.int32, device=dev)
    ctx.crc32(d_in, data.size, 0, res[0:1]); ctx.adler32(d_in, data.size, 1, res[1:2])
    torch.cuda.synchronize()
    r = res.cpu().numpy().view(np.uint32)
    assert (int(r[0]), int(r[1])) == (zo.port_crc32(data), zo.port_adler32(data))
    members, off = zo.gzip_members(data[: 64 * 4096], 4096, 1)
    d_m = torch.from_numpy(members).to(dev)
    d_io = torch.from_numpy(off.astype(np.int64)).to(dev); d_oo = torch.from_numpy(np.arange(65, dtype=np.int64) * 4096).to(dev)
    d_out = torch.zeros(64 * 4096, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(64, dtype=torch.int32, device=dev); checks = torch.zeros_like(sizes); status = torch.zeros_like(sizes)
    ctx.inflate_members(d_m, d_io, 64, 31, d_out, d_oo, sizes, checks, status, None, None)
    torch.cuda.synchronize()
    assert bool((status == 1).all()) and np.array_equal(d_out.cpu().numpy(), data[: 64 * 4096])
    print("K3 / K4 ok", flush=True)
if which in ("all", "host"):
    big = synth(33 << 20, seed=5)                              # >= 32 MiB: the streamed level-1 host path
    out = np.zeros(int(big.size * 1.2) + 65536, dtype=np.uint8)
    n_out, crc, _ = ctx.deflate_host(big, big.size, 65536, 1, False, out, out.size)
    import zlib
    assert zlib.decompress(out[:n_out].tobytes() + b"\x03\x00", wbits=-15) == big.tobytes() and crc == zlib.crc32(big.tobytes())
    print("streamed host path ok", flush=True)
ctx.close()
print("sanitize_smoke: all checks passed")
