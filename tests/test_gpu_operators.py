"""The operator surface (functable.h:26-42, deflate.h:121-131) one operator at a time, the way the reference's unit
tests exercise it: compare256 (test/test_compare256.cc: first mismatch at every index), chunkmemset_safe (byte-serial
overlap semantics), longest_match and insert_string (against the oracle's restatement), crc32 / adler32 through the
host-callable table."""
import ctypes
import zlib as pyzlib

import numpy as np
from synthdata import synth
import pytest

pytestmark = pytest.mark.gpu


def test_compare256_every_mismatch_index(pkg, ctx, zo):
    """test_compare256.cc:25-51: str1 all 'a'; str2 differs only at index i -> compare256 == i; identical -> 256."""
    import torch
    dev = f"cuda:{ctx.device}"
    n_pairs = 257
    stride = 288
    a = np.full((n_pairs, stride), ord("a"), dtype=np.uint8)
    b = a.copy()
    for i in range(256):
        b[i, i] = 0
    d_a = torch.zeros(n_pairs * stride + 64, dtype=torch.uint8, device=dev)
    d_b = torch.zeros(n_pairs * stride + 64, dtype=torch.uint8, device=dev)
    out = torch.zeros(n_pairs, dtype=torch.int32, device=dev)
    exp = np.array([zo.port().zo_compare256(a[i].ctypes.data, b[i].ctypes.data) for i in range(n_pairs)])
    assert np.array_equal(exp[:256], np.arange(256)) and exp[256] == 256
    for off_a in range(4):                                  # every relative alignment of the two operands
        for off_b in range(4):
            d_a.zero_(); d_b.zero_()
            d_a[off_a: off_a + n_pairs * stride] = torch.from_numpy(a.reshape(-1)).to(dev)
            d_b[off_b: off_b + n_pairs * stride] = torch.from_numpy(b.reshape(-1)).to(dev)
            ctx._check(pkg.lib().zng_b200_op_compare256(ctx._h, d_a.data_ptr() + off_a, d_b.data_ptr() + off_b, stride, n_pairs, out.data_ptr(), 0))
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy(), exp), (off_a, off_b)


def test_functable_host_table(pkg, ctx, zo):
    ft = pkg.functable()
    assert ft.chunksize() == 32
    s1 = np.full(256, ord("a"), dtype=np.uint8)
    for i in (0, 1, 7, 8, 9, 63, 64, 128, 200, 255):
        s2 = s1.copy(); s2[i] = 0
        assert ft.compare256(s1.ctypes.data, s2.ctypes.data) == i
    assert ft.compare256(s1.ctypes.data, s1.ctypes.data) == 256
    data = synth(300000, seed=2)
    assert ft.crc32(0, data.ctypes.data, data.size) == pyzlib.crc32(data.tobytes())
    assert ft.adler32(1, data.ctypes.data, data.size) == pyzlib.adler32(data.tobytes())
    # chunkmemset_safe(out, from, len, left): byte-serial overlap copy, clipped to `left`
    for dist, ln, left in ((1, 258, 300), (2, 100, 300), (3, 258, 258), (7, 33, 20), (40, 258, 258), (300, 258, 258)):
        buf = np.zeros(1000, dtype=np.uint8)
        buf[:300] = np.arange(300) % 251 + 1
        exp = buf.copy()
        for k in range(min(ln, left)):
            exp[300 + k] = exp[300 + k - dist]
        ret = ft.chunkmemset_safe(buf.ctypes.data + 300, buf.ctypes.data + 300 - dist, ln, left)
        assert ret == buf.ctypes.data + 300 + min(ln, left)
        assert np.array_equal(buf, exp), (dist, ln, left)


def test_chunkmemset_device(pkg, ctx):
    import torch
    dev = f"cuda:{ctx.device}"
    rng = np.random.default_rng(1)
    for dist in (1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 64, 100, 257, 258, 4000):
        for ln in (1, 3, 31, 32, 33, 64, 257, 258):
            base = rng.integers(0, 256, size=4096 + 600, dtype=np.uint8)
            exp = base.copy()
            for k in range(ln):
                exp[4096 + k] = exp[4096 + k - dist]
            d = torch.from_numpy(base).to(dev)
            ctx._check(pkg.lib().zng_b200_op_chunkmemset(ctx._h, d.data_ptr(), 4096, dist, ln, 0))
            torch.cuda.synchronize()
            assert np.array_equal(d.cpu().numpy(), exp), (dist, ln)


def test_insert_string_and_longest_match_vs_oracle(pkg, ctx, zo):
    import torch
    dev = f"cuda:{ctx.device}"
    rng = np.random.default_rng(3)
    makers = (lambda: synth(65536, seed=31)[:60000], lambda: rng.integers(0, 4, size=60000, dtype=np.uint8),
              lambda: np.tile(rng.integers(0, 256, size=37, dtype=np.uint8), 2000)[:60000])
    for trial, maker in enumerate(makers):
        data = np.ascontiguousarray(maker())
        n = data.size
        avail = n + 300
        win = np.zeros(avail, dtype=np.uint8); win[:n] = data
        if trial:
            win[n:] = rng.integers(0, 256, size=300, dtype=np.uint8)
        head = np.zeros(65536, dtype=np.uint16); prev = np.zeros(32768, dtype=np.uint16)
        d_win = torch.from_numpy(win).to(dev)
        d_head = torch.zeros(65536, dtype=torch.int16, device=dev); d_prev = torch.zeros(32768, dtype=torch.int16, device=dev)
        # insert_string over several ranges (one of them re-inserts an already inserted range)
        for s0, cnt in ((0, 1), (1, 5), (6, 31), (37, 32), (69, 33), (102, 20000), (50, 100), (20102, n - 4 - 20102)):
            zo.port().zo_insert_string(win.ctypes.data, avail, head.ctypes.data, prev.ctypes.data, s0, cnt)
            ctx._check(pkg.lib().zng_b200_op_insert_string(ctx._h, d_win.data_ptr(), d_head.data_ptr(), d_prev.data_ptr(), s0, cnt, 0))
        torch.cuda.synchronize()
        assert np.array_equal(d_head.cpu().numpy().view(np.uint16), head), trial
        assert np.array_equal(d_prev.cpu().numpy().view(np.uint16), prev), trial
        # longest_match queries: cur_match = the previous position with the same hash, incl. positions near the end
        pos = np.concatenate([rng.integers(33000, n - 4, size=1500), np.arange(n - 300, n - 3)]).astype(np.int64)
        cand = prev[pos & 32767].astype(np.int64)
        ok = (cand != 0) & (cand < pos) & (pos - cand <= 32506)
        pos, cand = pos[ok], cand[ok]
        assert pos.size > 100
        d_pos = torch.from_numpy(pos.astype(np.int32)).to(dev); d_cand = torch.from_numpy(cand.astype(np.int32)).to(dev)
        d_len = torch.zeros(pos.size, dtype=torch.int32, device=dev); d_start = torch.zeros_like(d_len)
        ctx._check(pkg.lib().zng_b200_op_longest_match(ctx._h, d_win.data_ptr(), n, d_prev.data_ptr(), d_pos.data_ptr(), d_cand.data_ptr(),
                                                      pos.size, d_len.data_ptr(), d_start.data_ptr(), 0))
        torch.cuda.synchronize()
        gl, gs = d_len.cpu().numpy(), d_start.cpu().numpy()
        st = ctypes.c_uint32(0)
        nmatch = 0
        for i in range(pos.size):
            el = zo.port().zo_longest_match_l2(win.ctypes.data, avail, n, prev.ctypes.data, int(pos[i]), int(cand[i]), ctypes.byref(st))
            if el >= 4:
                nmatch += 1
                assert gl[i] == el and gs[i] == st.value, (trial, i, int(pos[i]), int(cand[i]), gl[i], el, gs[i], st.value)
            else:
                assert gl[i] == 0, (trial, i)
        assert nmatch > 50
