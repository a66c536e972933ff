"""Pin the CPU restatement (oracle/paged_oracle.c) to the reference's own compiled code
(oracle/_ref, built from /root/reference by oracle/build_ref.sh) and to the committed golden
vectors.  CPU only."""
import os

import numpy as np
import pytest

import oracle_api as oa
from trace_driver import make_trace, run_trace

GEOMS = [(16, 12, 8), (4, 24, 6), (32, 100, 100), (8, 64, 16)]


def need_ref(bs, mb, mp, flavor="strict"):
    if not oa.have_ref(bs, mb, mp, flavor):
        pytest.skip(f"oracle/_ref variant bs{bs}_mb{mb}_mp{mp}_{flavor} not built (needs /root/reference)")


# ------------------------------------------------------------------ integer side
@pytest.mark.parametrize("bs,mb,mp", GEOMS)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_allocator_trace_bit_exact(bs, mb, mp, seed):
    need_ref(bs, mb, mp)
    ops = make_trace(seed, 600, mp, bs, n_active=min(mp, 10))
    ref = oa.RefManager(4, bs, mb, mp)
    orc = oa.OrcManager(4, bs, mb, mp)
    try:
        a = run_trace(ref, ops, mp, mb, snap_every=7)
        b = run_trace(orc, ops, mp, mb, snap_every=7)
        assert a == b
    finally:
        ref.close(); orc.close()


def test_survey_scripted_trace():
    """interleaved requests, free, refill, LRU whole-prompt eviction."""
    bs, mb, mp = 32, 100, 100
    need_ref(bs, mb, mp)
    for mk in (lambda: oa.RefManager(2, bs, mb, mp), lambda: oa.OrcManager(2, bs, mb, mp)):
        m = mk()
        try:
            for _ in range(3):
                for p in (0, 1, 2):
                    m.request_block(p)
            assert [m.table(p) for p in (0, 1, 2)] == [[0, 3, 6], [1, 4, 7], [2, 5, 8]]
            assert m.epoch() == 9
            m.free_blocks_for_prompt(1)
            for _ in range(4):
                m.request_block(3)
            assert m.table(3) == [1, 4, 7, 9]
            got = [m.request_block(4) for _ in range(95)]
            assert got[:90] == list(range(10, 100))
            assert got[90:] == [0, 3, 6, 2, 5]          # prompt 0 evicted first, then prompt 2
            assert m.table(0) == [] and m.table(2) == []
            assert m.request_block(100) == -1
        finally:
            m.close()


# ------------------------------------------------------------------ append + attention
def _fill_pages(mgr, prompt, kv, bs):
    """kv: (ntok, 2, C) -> request pages for `prompt` and copy rows in (test-side fill,
    like block_manager_test.c:8-29 writes through the KVBlock pointers)."""
    ntok = kv.shape[0]
    for t0 in range(0, ntok, bs):
        idx = mgr.request_block(prompt)
        k, v = mgr.page_arrays(idx)
        n = min(bs, ntok - t0)
        k[:n] = kv[t0:t0 + n, 0]
        v[:n] = kv[t0:t0 + n, 1]
        mgr.set_filled(idx, n)


@pytest.mark.parametrize("bs,mb,mp,T,C,NH,offset", [
    (2, 64, 8, 20, 10, 2, 0),        # test_paged_attn.c:184-188 shape
    (4, 24, 6, 13, 24, 3, 5),
    (8, 64, 16, 40, 64, 4, 17),
    (16, 100, 100, 64, 128, 2, 3),
    (32, 100, 100, 32, 768, 12, 18),  # paged_infer.c main: T=32, window slid by 18
])
def test_attention_full_window_bit_exact_strict(bs, mb, mp, T, C, NH, offset):
    need_ref(bs, mb, mp)
    ntok = T + offset
    kv = oa.normal((ntok, 2, C), seed=1234 + T)
    inp = oa.normal((1, T, 3 * C), seed=99 + C)
    ref = oa.RefManager(C, bs, mb, mp)
    orc = oa.OrcManager(C, bs, mb, mp)
    try:
        _fill_pages(ref, 0, kv, bs)
        _fill_pages(orc, 0, kv, bs)
        ra, ro, rp, rt = ref.attend(0, inp, 1, T, NH, offset, want_scratch=True)
        oa_, oo, op, ot = orc.attend(0, inp, 1, T, NH, offset, want_scratch=True)
        assert ra == oa_
        assert np.array_equal(ro.view(np.uint32), oo.view(np.uint32))
        # side outputs: only entries t2 <= t of preatt are written by the reference (:201)
        tri = np.tril(np.ones((T, T), dtype=bool))
        assert np.array_equal(rp[0][:, tri].view(np.uint32), op[0][:, tri].view(np.uint32))
        assert np.array_equal(rt.view(np.uint32), ot.view(np.uint32))
        # decode = last row of the window
        q_last = inp[0, T - 1, :C][None, :]
        dec = orc.decode_batch([0], NH, q_last, kv_start=[offset])
        assert np.array_equal(dec[0].view(np.uint32), ro[0, T - 1].view(np.uint32))
    finally:
        ref.close(); orc.close()


def test_fast_flavour_within_tolerance():
    """The Makefile's -Ofast build only defines the reference to ~3e-7."""
    bs, mb, mp, T, C, NH = 16, 100, 100, 256, 768, 12
    need_ref(bs, mb, mp, "fast")
    kv = oa.normal((T, 2, C), seed=1337)
    inp = oa.normal((1, T, 3 * C), seed=7)
    ref = oa.RefManager(C, bs, mb, mp, "fast")
    orc = oa.OrcManager(C, bs, mb, mp, "strict")
    try:
        _fill_pages(ref, 0, kv, bs)
        _fill_pages(orc, 0, kv, bs)
        _, ro = ref.attend(0, inp, 1, T, NH, 0)
        _, oo = orc.attend(0, inp, 1, T, NH, 0)
        err = np.abs(ro - oo).max() / np.abs(oo).max()
        assert err < 2e-6, err
    finally:
        ref.close(); orc.close()


def test_reference_differential_paged_equals_contiguous():
    """The one invariant the reference itself tests (test_paged_attn.c:244-248), on its own
    shape and value range U[0,100), here with a fixed seed and for the real function."""
    bs, mb, mp, T, C, NH = 2, 64, 8, 20, 10, 2
    need_ref(bs, mb, mp)
    import ctypes as Ct
    tl_path = os.path.join(oa.REF_DIR, "libref_train_strict.so")
    if not os.path.exists(tl_path):
        pytest.skip("libref_train not built")
    tl = Ct.CDLL(tl_path)
    inp = oa.uniform((1, T, 3 * C), 0.0, 100.0, seed=42)
    kv = np.stack([inp[0, :, C:2 * C], inp[0, :, 2 * C:]], axis=1)
    ref = oa.RefManager(C, bs, mb, mp)
    orc = oa.OrcManager(C, bs, mb, mp)
    try:
        _fill_pages(ref, 0, kv, bs)
        _fill_pages(orc, 0, kv, bs)
        _, ro = ref.attend(0, inp, 1, T, NH, 0)
        _, oo = orc.attend(0, inp, 1, T, NH, 0)
        out = np.zeros((1, T, C), dtype=np.float32)
        pre = np.zeros((1, NH, T, T), dtype=np.float32)
        att = np.zeros((1, NH, T, T), dtype=np.float32)
        tl.ref_attention_forward(oa.fptr(out), oa.fptr(pre), oa.fptr(att), oa.fptr(inp), 1, T, C, NH)
        assert np.abs(out - ro).max() <= 1e-2          # the reference's own tolerance (:8)
        assert np.array_equal(out.view(np.uint32), ro.view(np.uint32))   # in fact bit-identical
        assert np.array_equal(oo.view(np.uint32), ro.view(np.uint32))
    finally:
        ref.close(); orc.close()


def test_real_add_to_cache_sliding_window():
    """Drive the reference's real add_to_cache + collect_kv_blocks + attention_paged through
    main's pattern (paged_infer.c:1055-1057: T=32, first call n_tail=T, then n_tail=1 with
    offset 1..18) and require the restatement to agree bit-for-bit at every step."""
    bs, mb, mp, T, C, NH = 32, 100, 100, 32, 48, 4
    need_ref(bs, mb, mp)
    ref = oa.RefManager(C, bs, mb, mp)
    orc = oa.OrcManager(C, bs, mb, mp)
    try:
        stream = oa.normal((T + 18, 3 * C), seed=2024)
        for step in range(19):
            window = np.ascontiguousarray(stream[step:step + T][None])
            n_tail = T if step == 0 else 1
            ref.add_to_cache(window, 1, T, n_tail)
            assert orc.add_to_cache(window, 1, T, n_tail) >= 0
            assert ref.table(0) == orc.table(0)
            assert ref.epoch() == orc.epoch()
            _, ro = ref.attend(0, window, 1, T, NH, step)
            _, oo = orc.attend(0, window, 1, T, NH, step)
            assert np.array_equal(ro.view(np.uint32), oo.view(np.uint32)), step
        assert ref.table(0) == [0, 1]
        assert [ref.block_info(i)[0] for i in (0, 1)] == [32, 18]
        assert ref.epoch() == 19
        for idx in (0, 1):
            rk, rv = ref.page_arrays(idx)
            ok, ov = orc.page_arrays(idx)
            n = ref.block_info(idx)[0]
            assert np.array_equal(rk[:n], ok[:n]) and np.array_equal(rv[:n], ov[:n])
    finally:
        ref.close(); orc.close()


def test_rng_matches_reference():
    bs, mb, mp = 32, 100, 100
    need_ref(bs, mb, mp)
    import ctypes as Ct
    rl = oa.load_ref(bs, mb, mp)
    ol = oa.load_oracle()
    a, b = Ct.c_ulonglong(1337), Ct.c_ulonglong(1337)
    for _ in range(1000):
        assert rl.ref_random_u32(Ct.byref(a)) == ol.orc_random_u32(Ct.byref(b))
    assert rl.ref_random_f32(Ct.byref(a)) == ol.orc_random_f32(Ct.byref(b))


def test_matmul_cached_restatement():
    """Next-row checker: matmul_forward / matmul_cached, paged_infer.c:92-160."""
    bs, mb, mp = 32, 100, 100
    need_ref(bs, mb, mp)
    rl = oa.load_ref(bs, mb, mp)
    ol = oa.load_oracle()
    B, T, C = 2, 5, 24
    x = oa.normal((B, T, C), seed=5)
    w = oa.normal((3 * C, C), seed=6)
    bias = oa.normal((3 * C,), seed=7)
    for fn_r, fn_o in ((rl.ref_matmul_forward, ol.orc_matmul_forward), (rl.ref_matmul_cached, ol.orc_matmul_cached)):
        r = np.zeros((B, T, 3 * C), dtype=np.float32)
        o = np.zeros((B, T, 3 * C), dtype=np.float32)
        fn_r(oa.fptr(r), oa.fptr(x), oa.fptr(w), oa.fptr(bias), B, T, C, 3 * C)
        fn_o(oa.fptr(o), oa.fptr(x), oa.fptr(w), oa.fptr(bias), B, T, C, 3 * C)
        assert np.array_equal(r.view(np.uint32), o.view(np.uint32))


def test_layer_ops_restatement():
    """Next-row checkers: encoder / layernorm / gelu / residual / softmax / sample_mult
    restated in oracle/paged_oracle.c == the compiled reference, bit for bit (strict flavour)."""
    bs, mb, mp = 32, 100, 100
    need_ref(bs, mb, mp)
    rl = oa.load_ref(bs, mb, mp)
    ol = oa.load_oracle()
    B, T, C_, V = 2, 5, 48, 97
    x = oa.normal((B, T, C_), seed=11)
    w = oa.normal((C_,), seed=12)
    b = oa.normal((C_,), seed=13)
    # layernorm (+ mean, rstd)
    outs = []
    for lib, pre in ((rl, "ref_"), (ol, "orc_")):
        o = np.zeros_like(x); m = np.zeros((B, T), np.float32); r = np.zeros((B, T), np.float32)
        getattr(lib, pre + "layernorm_forward")(oa.fptr(o), oa.fptr(m), oa.fptr(r), oa.fptr(x), oa.fptr(w), oa.fptr(b), B, T, C_)
        outs.append((o, m, r))
    for a, c in zip(outs[0], outs[1]):
        assert np.array_equal(a.view(np.uint32), c.view(np.uint32))
    # gelu, residual
    y = oa.normal((B * T * C_,), seed=14) * np.float32(3.0)
    for name, args in (("gelu_forward", (y,)), ("residual_forward", (y, y[::-1].copy()))):
        res = []
        for lib, pre in ((rl, "ref_"), (ol, "orc_")):
            o = np.zeros_like(y)
            getattr(lib, pre + name)(oa.fptr(o), *[oa.fptr(a) for a in args], y.size)
            res.append(o)
        assert np.array_equal(res[0].view(np.uint32), res[1].view(np.uint32)), name
    # encoder
    wte = oa.normal((V, C_), seed=15); wpe = oa.normal((T, C_), seed=16)
    tok = (np.arange(B * T, dtype=np.int32) * 7 % V).astype(np.int32)
    res = []
    for lib, pre in ((rl, "ref_"), (ol, "orc_")):
        o = np.zeros((B, T, C_), np.float32)
        getattr(lib, pre + "encoder_forward")(oa.fptr(o), oa.iptr(tok), oa.fptr(wte), oa.fptr(wpe), B, T, C_)
        res.append(o)
    assert np.array_equal(res[0], res[1])
    # softmax + sample_mult over a range of coins
    logits = oa.normal((B, T, V), seed=17) * np.float32(4.0)
    res = []
    for lib, pre in ((rl, "ref_"), (ol, "orc_")):
        pr = np.zeros_like(logits)
        getattr(lib, pre + "softmax_forward")(oa.fptr(pr), oa.fptr(logits), B, T, V)
        res.append(pr)
    assert np.array_equal(res[0].view(np.uint32), res[1].view(np.uint32))
    row = np.ascontiguousarray(res[0][0, 0])
    for coin in (0.0, 1e-7, 0.25, 0.5, 0.75, 0.999999, 0.9999999):
        assert rl.ref_sample_mult(oa.fptr(row), V, coin) == ol.orc_sample_mult(oa.fptr(row), V, coin)
