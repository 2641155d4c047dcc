"""Deterministic structured fuzz through the GPU path: the inputs of tests/test_oracle_deflate.py::test_port_matches_reference_on_
structured_random_inputs (small alphabets, runs, copies at distances around MAX_DIST and the window size), every level, one
fresh Z_FINISH chunk per input, plus primed chunks at level 1 -- byte for byte against the oracle (which that test pins to the
unmodified reference on the same inputs)."""
import numpy as np
import pytest

from test_oracle_deflate import _fuzz_input

pytestmark = pytest.mark.gpu


def test_structured_random_inputs_every_level(pkg, ctx, zo):
    import torch
    rng = np.random.default_rng(2026)
    sizes = [int(x) for x in rng.integers(0, 3000, size=80)] + [int(x) for x in rng.integers(60000, 65537, size=30)] + [65536] * 10
    stride = pkg.deflate_bound(65536)
    dev = f"cuda:{ctx.device}"
    slots = torch.zeros(stride, dtype=torch.uint8, device=dev)
    sizes_d = torch.zeros(1, dtype=torch.int32, device=dev)
    checked = 0
    for n in sizes:
        d = _fuzz_input(rng, n)
        d_in = torch.from_numpy(d).to(dev) if n else torch.zeros(16, dtype=torch.uint8, device=dev)
        for level in (1, 2, 3, 4, 5, 6):
            slots.fill_(0xEE)
            ctx.deflate_chunks(d_in, n, 65536, level, pkg.Z_FINISH, slots, stride, sizes_d, None, None)
            torch.cuda.synchronize()
            exp, es, _, _ = zo.port_deflate_chunks(d, 65536, level, 4, stride, nthreads=1)
            if n == 0:
                continue                                      # no chunk, nothing written
            got = int(sizes_d.cpu().numpy().view(np.uint32)[0])
            assert got == int(es[0]), (n, level, got, int(es[0]))
            assert np.array_equal(slots[:got].cpu().numpy(), exp[0, :got]), (n, level)
            checked += 1
    assert checked >= 700
    # the same generator, primed level-1 chunks
    for n in sizes[-25:]:
        d = _fuzz_input(rng, 65536 + n)
        d_in = torch.from_numpy(d).to(dev)
        out, stride2, sz, crcs, adlers = ctx.alloc_chunk_outputs(d.size, 65536, adler=True)
        ctx.deflate_chunks_primed(d_in, d.size, 65536, 1, 2, out, stride2, sz, crcs, adlers)
        torch.cuda.synchronize()
        exp, es, _, _ = zo.port_deflate_chunks_primed(d, 65536, 1, 2, stride2, nthreads=1)
        got = sz.cpu().numpy().view(np.uint32)[: len(es)]
        host = out.cpu().numpy().reshape(-1, stride2)
        assert np.array_equal(got, es), ("primed", n)
        assert all(np.array_equal(host[i, : es[i]], exp[i, : es[i]]) for i in range(len(es))), ("primed", n)
