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


# ---- every slot of the 15-slot table against the reference's OWN dispatched variant (oracle/_ref + oracle/ref_ops.c) ------------
def _need_ref(zo):
    if not zo.have_ref() or not hasattr(zo.ref(), "refops_longest_match"):
        pytest.skip("oracle/_ref with ref_ops.c not built")
    return zo.ref()


def test_table_has_the_references_15_slots_in_order(pkg):
    names = [f[0] for f in pkg.Functable._fields_]
    assert names == ["force_init", "adler32", "adler32_fold_copy", "chunkmemset_safe", "chunksize", "compare256", "crc32", "crc32_fold",
                     "crc32_fold_copy", "crc32_fold_final", "crc32_fold_reset", "inflate_fast", "longest_match", "longest_match_slow",
                     "slide_hash"]                       # functable.h:26-42
    ft = pkg.functable()
    ft.force_init()
    for n in names:
        assert ctypes.cast(getattr(ft, n), ctypes.c_void_p).value, n


def test_fold_slots_vs_reference(pkg, ctx, zo):
    R = _need_ref(zo)
    ft = pkg.functable()
    data = synth(700001, seed=12)
    for piece in (700001, 65536, 4097, 333):
        f = pkg.Crc32Fold()
        assert ft.crc32_fold_reset(ctypes.byref(f)) == 0
        dst = np.zeros(data.size, dtype=np.uint8)
        for o in range(0, data.size, piece):
            k = min(piece, data.size - o)
            if (o // piece) % 2:
                ft.crc32_fold(ctypes.byref(f), data.ctypes.data + o, k, 0)
                dst[o:o + k] = data[o:o + k]
            else:
                ft.crc32_fold_copy(ctypes.byref(f), dst.ctypes.data + o, data.ctypes.data + o, k)
        rdst = np.zeros(data.size, dtype=np.uint8)
        assert ft.crc32_fold_final(ctypes.byref(f)) == R.refops_crc32_fold(data.ctypes.data, data.size, piece, rdst.ctypes.data)
        assert np.array_equal(dst, data) and np.array_equal(rdst, data)
    for init in (1, 0x12345678):
        dst = np.zeros(data.size, dtype=np.uint8); rdst = np.zeros(data.size, dtype=np.uint8)
        got = ft.adler32_fold_copy(init, dst.ctypes.data, data.ctypes.data, data.size)
        assert got == R.refops_adler32_fold_copy(init, rdst.ctypes.data, data.ctypes.data, data.size)
        assert np.array_equal(dst, data)


def test_compare256_chunkmemset_update_hash_vs_reference(pkg, ctx, zo):
    R = _need_ref(zo)
    ft = pkg.functable()
    rng = np.random.default_rng(9)
    for _ in range(40):
        a = rng.integers(0, 3, size=256, dtype=np.uint8); b = a.copy()
        k = int(rng.integers(0, 300))
        if k < 256:
            b[k] ^= 1
        assert ft.compare256(a.ctypes.data, b.ctypes.data) == R.refops_compare256(a.ctypes.data, b.ctypes.data)
    for dist, ln, left in ((1, 258, 300), (2, 100, 300), (3, 258, 258), (7, 33, 20), (40, 258, 258), (300, 258, 258), (5, 1, 1), (16, 64, 64)):
        base = np.zeros(1200, dtype=np.uint8); base[:400] = rng.integers(1, 256, size=400, dtype=np.uint8)
        g = base.copy(); r = base.copy()
        adv = R.refops_chunkmemset_safe(r.ctypes.data, 400, dist, ln, left)
        ret = ft.chunkmemset_safe(g.ctypes.data + 400, g.ctypes.data + 400 - dist, ln, left)
        assert ret - (g.ctypes.data + 400) == adv
        assert np.array_equal(g[:400 + adv], r[:400 + adv]), (dist, ln, left)       # the reference may scribble past `len` inside `left`
    for v in (0, 1, 0x61626364, 0xffffffff, 0x9e3779b9):
        assert pkg.lib().zng_b200_update_hash(0, v) == R.refops_update_hash(0, v)


def test_insert_string_quick_insert_slide_hash_vs_reference(pkg, ctx, zo):
    R = _need_ref(zo)
    ft = pkg.functable()
    rng = np.random.default_rng(4)
    data = np.ascontiguousarray(synth(65536, seed=77)[:50000])
    head = np.zeros(65536, dtype=np.uint16); prev = np.zeros(32768, dtype=np.uint16)
    rhead = head.copy(); rprev = prev.copy()
    st = pkg.MatchState(data.ctypes.data, data.size, head.ctypes.data, prev.ctypes.data, 0, 0, 0, 2)
    for s0, cnt in ((0, 1), (1, 40), (41, 3000), (20, 64), (3041, 30000)):
        pkg.lib().zng_b200_insert_string(ctypes.byref(st), s0, cnt)
        R.refops_insert_string(data.ctypes.data, data.size, rhead.ctypes.data, rprev.ctypes.data, s0, cnt)
        assert np.array_equal(head, rhead) and np.array_equal(prev, rprev), (s0, cnt)
    for s0 in (33041, 33042, 100, 40000, 40000):
        got = pkg.lib().zng_b200_quick_insert_string(ctypes.byref(st), s0)
        exp = R.refops_insert_string(data.ctypes.data, data.size, rhead.ctypes.data, rprev.ctypes.data, s0, 0)
        assert got == exp and np.array_equal(head, rhead) and np.array_equal(prev, rprev), s0
    # slide_hash on the populated tables plus random garbage
    head[rng.integers(0, 65536, size=3000)] = rng.integers(0, 65536, size=3000).astype(np.uint16)
    rhead = head.copy(); rprev = prev.copy()
    ft.slide_hash(ctypes.byref(st))
    assert R.refops_slide_hash(rhead.ctypes.data, rprev.ctypes.data) == 0
    assert np.array_equal(head, rhead) and np.array_equal(prev, rprev)


@pytest.mark.parametrize("level", [2, 3, 4, 5, 6])
def test_longest_match_slot_vs_reference_every_level(pkg, ctx, zo, level):
    """functable.longest_match through the table, one query per call, against the reference's dispatched variant with
    configuration_table[level]'s chain / nice parameters; the batched device form on the same queries as well."""
    import torch
    R = _need_ref(zo)
    ft = pkg.functable()
    rng = np.random.default_rng(level)
    dev = f"cuda:{ctx.device}"
    for trial, data in enumerate((synth(65536, seed=31)[:40000], rng.integers(0, 3, size=40000, dtype=np.uint8),
                                  np.tile(rng.integers(0, 256, size=97, dtype=np.uint8), 500)[:40000])):
        data = np.ascontiguousarray(data)
        n = data.size
        head = np.zeros(65536, dtype=np.uint16); prev = np.zeros(32768, dtype=np.uint16)
        R.refops_insert_string(data.ctypes.data, n, head.ctypes.data, prev.ctypes.data, 0, n - 4)
        pos = np.concatenate([rng.integers(33000, n - 4, size=24), np.arange(n - 12, n - 3)]).astype(np.int64)
        cand = prev[pos & 32767].astype(np.int64)
        ok = (cand != 0) & (cand < pos) & (pos - cand <= 32506)
        pos, cand = pos[ok], cand[ok]
        exp = []
        ms = ctypes.c_uint32(0)
        for p_, c_ in zip(pos, cand):
            el = R.refops_longest_match(data.ctypes.data, n, prev.ctypes.data, int(p_), int(c_), level, ctypes.byref(ms))
            exp.append((el, ms.value))
        st = pkg.MatchState(data.ctypes.data, n, head.ctypes.data, prev.ctypes.data, 0, 0, 0, level)
        for (p_, c_), (el, es) in list(zip(zip(pos, cand), exp))[:12]:
            st.strstart = int(p_); st.lookahead = n - int(p_); st.match_start = 0xffffffff
            gl = ft.longest_match(ctypes.byref(st), int(c_))
            assert gl == el, (level, trial, int(p_), int(c_), gl, el)
            if el >= 4:
                assert st.match_start == es
        # batched device form
        win = np.zeros(n + 600, dtype=np.uint8); win[:n] = data
        d_win = torch.from_numpy(win).to(dev); d_prev = torch.from_numpy(prev.view(np.int16)).to(dev)
        d_pos = torch.from_numpy(pos.astype(np.int32)).to(dev); d_cand = torch.from_numpy(cand.astype(np.int32)).to(dev)
        d_len = torch.zeros(pos.size, dtype=torch.int32, device=dev); d_start = torch.zeros_like(d_len)
        ctx._check(pkg.lib().zng_b200_op_longest_match_level(ctx._h, d_win.data_ptr(), n, d_prev.data_ptr(), d_pos.data_ptr(), d_cand.data_ptr(),
                                                            pos.size, level, d_len.data_ptr(), d_start.data_ptr(), 0))
        torch.cuda.synchronize()
        gl, gs = d_len.cpu().numpy(), d_start.cpu().numpy()
        for i, (el, es) in enumerate(exp):
            if el >= 4:
                assert gl[i] == el and gs[i] == es, (level, trial, i, int(pos[i]), int(cand[i]), gl[i], el)
            else:
                assert gl[i] == 0
    assert ft.longest_match_slow(ctypes.byref(st), 1) == 0          # levels 7-9: slot present, outside the hot path
