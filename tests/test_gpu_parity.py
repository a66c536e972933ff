"""-m gpu parity tests proper: every call goes through the C ABI of libpaged_attn.so; the checker
is the CPU oracle (pinned to the compiled reference by test_oracle_pinned.py) and, where the
reference itself was built (oracle/_ref), the reference's own attention_paged."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_api as oa
from gpu_common import assert_close_tc, assert_close_tc3, Scenario, assert_close, assert_close_illconditioned, pa

pytestmark = pytest.mark.gpu


# --------------------------------------------------------------------------------- KV append
@pytest.mark.parametrize("NH,hs,bs", [(12, 64, 16), (25, 64, 16), (4, 128, 32), (2, 5, 2)])
def test_append_bit_exact(NH, hs, bs):
    Cc = NH * hs
    B = 7
    eng = pa.PagedAttn(bs, 64, B, NH, hs, n_layers=2, max_batch_tokens=256)
    orc = oa.OrcManager(Cc, bs, 64, B)
    try:
        rng = np.random.default_rng(0)
        for step in range(6):
            seqs = rng.permutation(B)[: int(rng.integers(1, B + 1))].astype(np.int32)
            n_new = (rng.integers(1, 2 * bs, size=len(seqs)) if step == 0 else np.ones(len(seqs))).astype(np.int32)
            ntok = int(n_new.sum())
            qkv = oa.normal((ntok, 3 * Cc), seed=10 + step)
            assert eng.step_begin(seqs, n_new) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            d = pa.DevBuf.from_numpy(qkv)
            pa.check(eng.append(1, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc), "append")
            eng.sync()
            slots = eng.slot_mapping()
            k, v = eng.read_pool_rows(1, slots)
            assert np.array_equal(k.view(np.uint32), qkv[:, Cc:2 * Cc].view(np.uint32))
            assert np.array_equal(v.view(np.uint32), qkv[:, 2 * Cc:].view(np.uint32))
            # same bytes at the same logical positions as the oracle's add_to_cache
            row = 0
            for s, n in zip(seqs, n_new):
                for _ in range(int(n)):
                    orc.add_to_cache(qkv[row][None, None, :], 1, 1, 1, prompt=int(s))
                    row += 1
            for s in seqs:
                L = orc.context_len(int(s))
                assert L == eng.seq_len(int(s))
                pos = L - 1
                ok, ov = orc.page_arrays(orc.table(int(s))[pos // bs])
                gk, gv = eng.read_pool_rows(1, [eng.table(int(s))[pos // bs] * bs + pos % bs])
                assert np.array_equal(gk[0], ok[pos % bs]) and np.array_equal(gv[0], ov[pos % bs])
            d.free()
        # layer 0 untouched
        k0, _ = eng.read_pool_rows(0, eng.slot_mapping())
        assert not k0.any()
    finally:
        eng.close(); orc.close()


# --------------------------------------------------------------------------------- decode
RAGGED = [1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 49, 100, 255, 256, 257, 333]

DECODE_CASES = [
    # NH, hs, bs, ctx, shuffle
    pytest.param(12, 64, 16, [256], False, id="cfg1-b1-ctx256"),
    pytest.param(12, 64, 16, RAGGED, True, id="ragged-bs16-shuffled"),
    pytest.param(12, 64, 32, RAGGED, False, id="ragged-bs32-refdefault"),
    pytest.param(12, 64, 8, RAGGED, True, id="ragged-bs8"),
    pytest.param(12, 64, 4, [1, 3, 4, 5, 64, 65], True, id="bs4"),
    pytest.param(25, 64, 16, [128, 1024, 77, 513], True, id="xl-25heads"),
    pytest.param(4, 128, 16, [1, 17, 300, 2048], True, id="hs128"),
    pytest.param(4, 128, 32, [5, 64, 700], False, id="hs128-bs32"),
    pytest.param(8, 128, 8, [9, 130], True, id="hs128-bs8"),
    pytest.param(12, 64, 16, [0, 5, 0, 40], False, id="empty-sequences"),
    pytest.param(2, 64, 16, [1 + (i * 7) % 40 for i in range(1100)], True, id="batch-1100-beyond-smem-prefix-table"),
    pytest.param(4, 128, 16, [32768, 5], True, id="cfg5-32k-context"),
]


@pytest.mark.parametrize("NH,hs,bs,ctx,shuffle", DECODE_CASES)
@pytest.mark.parametrize("path", [1, 2, 3], ids=["stream", "generic", "small-batch"])
def test_decode_matches_oracle(NH, hs, bs, ctx, shuffle, path):
    sc = Scenario(NH, hs, bs, ctx, shuffle=shuffle, seed=77)
    try:
        q = oa.normal((sc.B, sc.C), seed=5)
        want = sc.oracle_decode(q)
        got = sc.decode(q, path=path)
        assert_close(got, want, f"decode path={path}")
    finally:
        sc.close()


def test_decode_small_batch_kernel_head_dim_32():
    """head_dim 32 (8 lanes per token, four tokens per warp step): outside the stream kernel's domain, inside
    the small-batch kernel's -- automatic choice, forced small-batch kernel and generic kernel agree with the oracle."""
    sc = Scenario(4, 32, 16, [5, 100, 33, 1, 64, 257], shuffle=True, seed=21)
    try:
        q = oa.normal((sc.B, sc.C), seed=22)
        want = sc.oracle_decode(q)
        for path in (0, 3, 2):
            assert_close(sc.decode(q, path=path), want, f"hs32 path={path}")
    finally:
        sc.close()


@pytest.mark.parametrize("hpg,stages,grid", [(1, 0, 0), (2, 2, 0), (3, 3, 7), (4, 0, 1), (6, 0, 0), (12, 2, 0),
                                             (12, 4, 148), (12, 0, 1184), (0, 0, 5)])
def test_decode_tile_shapes_and_splits(hpg, stages, grid):
    """Every tile shape / ring depth / split count must give the same answer (split partials are
    merged in-kernel by the last-arriving CTA)."""
    sc = Scenario(12, 64, 16, [700, 3, 129, 64, 1000, 17], shuffle=True, seed=3)
    try:
        q = oa.normal((sc.B, sc.C), seed=6)
        want = sc.oracle_decode(q)
        for rep in range(3):      # repeated launches: the arrival counters must self-reset
            got = sc.decode(q, path=1, hpg=hpg, stages=stages, grid=grid)
            assert_close(got, want, f"hpg={hpg} stages={stages} grid={grid} rep={rep}")
    finally:
        sc.close()


@pytest.mark.parametrize("static_pct,dyn_units,grid", [(75, 2, 0), (50, 1, 0), (1, 1, 0), (90, 7, 0), (100, 0, 0),
                                                       (60, 3, 37), (30, 2, 300)])
def test_decode_dynamic_ranges(static_pct, dyn_units, grid):
    """Static head + dynamically claimed tail of the page stream: any split must give the same
    answer, launch after launch (the scheduler words and arrival counters reset themselves)."""
    sc = Scenario(12, 64, 16, [1024, 900, 3, 129, 64, 1000, 17, 512, 700, 333, 1, 256], shuffle=True, seed=13)
    try:
        q = oa.normal((sc.B, sc.C), seed=6)
        want = sc.oracle_decode(q)
        for rep in range(3):
            got = sc.decode(q, path=1, static_pct=static_pct, dyn_units=dyn_units, grid=grid)
            assert_close(got, want, f"static_pct={static_pct} dyn_units={dyn_units} grid={grid} rep={rep}")
    finally:
        sc.close()


@pytest.mark.parametrize("path", [0, 1, 3], ids=["auto", "stream", "small-batch"])
@pytest.mark.parametrize("NH,hs,bs,hpg", [(12, 64, 16, 0), (12, 64, 16, 3), (25, 64, 16, 0), (4, 128, 32, 0), (12, 64, 8, 4),
                                          (2, 5, 2, 0)])
def test_decode_append_fused(NH, hs, bs, hpg, path):
    """pa_decode_append = pa_append + pa_decode in one launch: the new token's K/V row is read
    from the step's k/v rows, used, and stored to its slot (bit-exact copy); next step sees it.
    Stream kernel, small-batch kernel (one CTA per sequence and head) and the automatic choice
    (head_dim 5: append kernel + generic rows kernel)."""
    if path != 0 and hs not in (64, 128):
        pytest.skip("stream / small-batch kernels: head_dim 64/128")
    Cc = NH * hs
    ctx0 = [0, 1, 15, 16, 17, 31, 32, 100, 255, 256, 700]      # page boundaries on both sides
    B = len(ctx0)
    sc = Scenario(NH, hs, bs, ctx0, n_layers=2, layer=1, seed=71, shuffle=True, extra_blocks=64)
    try:
        eng, orc = sc.eng, sc.orc
        eng.tune(pa.PA_TUNE_HEADS_PER_TILE, hpg)
        eng.tune(pa.PA_TUNE_DECODE_PATH, path)
        for step in range(4):
            qkv = oa.normal((B, 3 * Cc), seed=300 + step)
            assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            for s in range(B):
                orc.add_to_cache(qkv[s][None, None, :], 1, 1, 1, prompt=s)
            want = orc.decode_batch(sc.seq_ids, NH, qkv[:, :Cc])
            d = pa.DevBuf.from_numpy(qkv)
            o = pa.DevBuf(B * Cc * 4)
            pa.check(eng.decode_append(1, d.ptr, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc, o.ptr, Cc), "decode_append")
            eng.sync()
            assert_close(o.download((B, Cc)), want, f"fused step {step}")
            k, v = eng.read_pool_rows(1, eng.slot_mapping())
            assert np.array_equal(k.view(np.uint32), qkv[:, Cc:2 * Cc].view(np.uint32))
            assert np.array_equal(v.view(np.uint32), qkv[:, 2 * Cc:].view(np.uint32))
            d.free(); o.free()
        # a step with more than one new token per sequence cannot be fused
        assert eng.step_begin([0], [2]) == 0
        pa.check(eng.upload(), "upload")
        assert eng.decode_append(1, 1, 1, 1, 3 * Cc, 1, Cc) == pa.PA_ERR_INVALID
    finally:
        sc.close()


@pytest.mark.parametrize("hpg", [0, 8, 4, 1], ids=["auto", "hpg8", "hpg4", "hpg1"])
def test_decode_cfg5_shape_column_slices(hpg):
    """BASELINE configs[4] at the shape bench.py times (`long`): 32 heads x head_dim 128 (C = 4096), block 16,
    32k-token context.  A whole page (256 KiB) does not fit the ring, so the stream kernel's tile is a
    COLUMN SLICE of hpg heads, one bulk copy per row (pa_kernels.cu, the `hpg < NH` path at head_dim 128);
    the plan bench.py reports is hpg 8.  Contexts: 32768 (2048 pages), 5 (one partial page), 4097 (one
    token into page 257).  Read-only decode and the fused append, launch after launch."""
    NH, hs, bs = 32, 128, 16
    Cc = NH * hs
    ctx = [32768, 5, 4097]
    sc = Scenario(NH, hs, bs, ctx, shuffle=True, seed=55, extra_blocks=16)
    try:
        q = oa.normal((sc.B, sc.C), seed=56)
        want = sc.oracle_decode(q)
        for rep in range(2):
            got = sc.decode(q, path=1, hpg=hpg)
            assert_close(got, want, f"cfg5 shape hpg={hpg} rep={rep}")
        if hpg:
            assert sc.eng.lib.pa_tune_get(sc.eng.h, pa.PA_TUNE_LAST_HPG) == hpg
        else:
            assert sc.eng.lib.pa_tune_get(sc.eng.h, pa.PA_TUNE_LAST_HPG) in (1, 2, 4, 8), "auto plan is a column slice at C=4096"
        # fused append: contexts grow to 32769 / 6 / 4098 (the first crosses into a new page)
        eng, orc = sc.eng, sc.orc
        B = sc.B
        for step in range(2):
            qkv = oa.normal((B, 3 * Cc), seed=400 + step)
            assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            for s_ in range(B):
                orc.add_to_cache(qkv[s_][None, None, :], 1, 1, 1, prompt=s_)
            want = orc.decode_batch(sc.seq_ids, NH, qkv[:, :Cc])
            d = pa.DevBuf.from_numpy(qkv)
            o = pa.DevBuf(B * Cc * 4)
            pa.check(eng.decode_append(0, d.ptr, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc, o.ptr, Cc), "decode_append")
            eng.sync()
            assert_close(o.download((B, Cc)), want, f"cfg5 shape fused hpg={hpg} step {step}")
            k, v = eng.read_pool_rows(0, eng.slot_mapping())
            assert np.array_equal(k.view(np.uint32), qkv[:, Cc:2 * Cc].view(np.uint32))
            assert np.array_equal(v.view(np.uint32), qkv[:, 2 * Cc:].view(np.uint32))
            d.free(); o.free()
    finally:
        sc.close()


def test_decode_sliding_window():
    """Reference `offset` (paged_infer.c:190,1057): row attends cached tokens [kv_start, ctx)."""
    sc = Scenario(12, 64, 16, [100, 64, 33, 500], shuffle=True, seed=9)
    try:
        q = oa.normal((sc.B, sc.C), seed=2)
        for kv_start in ([0, 0, 0, 0], [1, 16, 32, 17], [99, 63, 0, 255], [37, 5, 31, 499]):
            want = sc.oracle_decode(q, kv_start=kv_start)
            for path in (1, 2, 3):
                got = sc.decode(q, kv_start=kv_start, path=path)
                assert_close(got, want, f"window {kv_start} path={path}")
    finally:
        sc.close()


def test_decode_large_logits_reference_value_range():
    """U[0,100) inputs as in test_paged_attn.c:201 -> logits of order 1e5: softmax is one-hot-ish
    and the -10000 initial maximum never wins."""
    sc = Scenario(2, 64, 16, [20, 70], seed=11, dist="uniform")
    try:
        q = oa.uniform((sc.B, sc.C), 0.0, 100.0, seed=4)
        want, truth = sc.oracle_decode(q), sc.oracle_decode_f64(q)
        for path in (1, 2, 3):
            assert_close_illconditioned(sc.decode(q, path=path), want, truth, f"large logits path={path}")
    finally:
        sc.close()


def test_decode_minus_10000_floor():
    """Scores far below the reference's initial maximum (-10000.0f, paged_infer.c:187) underflow
    to exp()=0 and the 0-sum guard (:213) yields zeros -- reproduce, do not 'fix'."""
    sc = Scenario(2, 64, 16, [5, 40], seed=12)
    try:
        q = np.full((sc.B, sc.C), -3000.0, dtype=np.float32)
        # make every key strongly positive so q.k*scale << -10000
        sc.pool_k[:] = 10.0
        lib = sc.eng.lib
        pa.check(lib.pa_memcpy_h2d(sc.eng.pool_k(0), sc.pool_k.ctypes.data, sc.pool_k.nbytes, None), "h2d")
        for s in range(sc.B):
            for idx in sc.orc.table(s):
                k, _ = sc.orc.page_arrays(idx)
                k[:] = 10.0
        want = sc.oracle_decode(q)
        assert not want.any()
        for path in (1, 2):
            got = sc.decode(q, path=path)
            assert np.array_equal(got, want)
    finally:
        sc.close()


def test_decode_generic_odd_shapes():
    """The reference test's own tiny shape (test_paged_attn.c:184-188: C=10, NH=2, block 2) and
    other shapes outside the stream kernel's domain go to the generic SIMT kernel."""
    for NH, hs, bs, ctx in [(2, 5, 2, [20, 1, 7]), (3, 48, 4, [33, 9]), (1, 200, 16, [50]), (5, 96, 32, [70, 31])]:
        sc = Scenario(NH, hs, bs, ctx, shuffle=True, seed=21)
        try:
            q = oa.normal((sc.B, sc.C), seed=8)
            assert_close(sc.decode(q, path=0), sc.oracle_decode(q), f"generic NH={NH} hs={hs} bs={bs}")
            assert sc.eng.decode(0, None, sc.C, None, sc.C) == pa.PA_ERR_INVALID
        finally:
            sc.close()


def test_decode_physical_placement_is_irrelevant():
    """Size-independent property: where pages live in the pool (block-table permutation) must not
    change a single bit -- the work split depends on logical pages only."""
    ctx = [513, 64, 1000, 31]
    a = Scenario(12, 64, 16, ctx, shuffle=False, seed=31)
    b = Scenario(12, 64, 16, ctx, shuffle=True, seed=31)
    try:
        # same logical content in both: copy a's logical rows into b's physical slots
        for s in range(len(ctx)):
            ta, tb = a.eng.table(s), b.eng.table(s)
            for ja, jb in zip(ta, tb):
                b.pool_k[jb * 16:(jb + 1) * 16] = a.pool_k[ja * 16:(ja + 1) * 16]
                b.pool_v[jb * 16:(jb + 1) * 16] = a.pool_v[ja * 16:(ja + 1) * 16]
        lib = b.eng.lib
        pa.check(lib.pa_memcpy_h2d(b.eng.pool_k(0), b.pool_k.ctypes.data, b.pool_k.nbytes, None), "h2d")
        pa.check(lib.pa_memcpy_h2d(b.eng.pool_v(0), b.pool_v.ctypes.data, b.pool_v.nbytes, None), "h2d")
        q = oa.normal((len(ctx), 768), seed=1)
        ga, gb = a.decode(q, path=1), b.decode(q, path=1)
        assert np.array_equal(ga.view(np.uint32), gb.view(np.uint32))
    finally:
        a.close(); b.close()


def test_decode_constant_values_property():
    """softmax weights sum to 1: with every V row equal to c the output is c."""
    sc = Scenario(12, 64, 16, [1, 100, 1024], seed=41)
    try:
        c = oa.normal((sc.C,), seed=99)
        sc.pool_v[:] = c
        pa.check(sc.eng.lib.pa_memcpy_h2d(sc.eng.pool_v(0), sc.pool_v.ctypes.data, sc.pool_v.nbytes, None), "h2d")
        q = oa.normal((sc.B, sc.C), seed=3)
        got = sc.decode(q, path=1)
        assert np.abs(got - c[None, :]).max() <= 2e-6 * np.abs(c).max() + 1e-6
    finally:
        sc.close()


# --------------------------------------------------------------------------------- BASELINE full sizes
def _ctx_cfg3():
    rng = np.random.default_rng(42)
    return [int(x) for x in rng.integers(128, 1025, size=256)]


@pytest.mark.parametrize("name,NH,hs,ctx", [
    ("cfg2: 64 sequences x 1024 ctx, GPT-2 small", 12, 64, [1024] * 64),
    ("cfg3: batch 256, mixed ctx 128..1024", 12, 64, _ctx_cfg3()),
    ("cfg4: GPT-2 XL shape, one GPU's 64 sequences x 1024", 25, 64, [1024] * 64),
], ids=["cfg2", "cfg3", "cfg4-xl"])
def test_baseline_full_size_decode_parity_and_properties(name, NH, hs, ctx):
    """BASELINE.json configurations at their full single-GPU sizes, fragmented block tables: direct
    parity with the oracle (the last-row restatement is cheap enough), plus the size-independent
    properties of the domain: constant values come back unchanged (weights sum to 1), the output is
    linear in V, and the result does not depend on how the page stream is split over CTAs."""
    sc = Scenario(NH, hs, 16, ctx, shuffle=True, seed=2024)
    try:
        q = oa.normal((sc.B, sc.C), seed=8)
        want = sc.oracle_decode(q)
        got = sc.decode(q, path=1)
        assert_close(got, want, name)
        # split invariance: another tile shape / grid gives the same numbers up to summation order
        other = sc.decode(q, path=1, hpg=1 if NH % 2 else 2, grid=37)
        assert np.abs(other - got).max() <= 2e-6 * np.abs(got).max() + 1e-6
        lib = sc.eng.lib
        v1 = sc.pool_v
        # linearity in V: attention(V1 + V2) = attention(V1) + attention(V2)
        v2 = oa.normal(v1.shape, seed=77)
        pa.check(lib.pa_memcpy_h2d(sc.eng.pool_v(0), v2.ctypes.data, v2.nbytes, None), "h2d")
        got2 = sc.decode(q, path=1)
        vs = (v1 + v2).astype(np.float32)
        pa.check(lib.pa_memcpy_h2d(sc.eng.pool_v(0), vs.ctypes.data, vs.nbytes, None), "h2d")
        got12 = sc.decode(q, path=1)
        assert np.abs(got12 - (got + got2)).max() <= 1e-5 * np.abs(got12).max()
        # constant V rows -> the constant
        c = oa.normal((sc.C,), seed=99)
        vs[:] = c
        pa.check(lib.pa_memcpy_h2d(sc.eng.pool_v(0), vs.ctypes.data, vs.nbytes, None), "h2d")
        gc = sc.decode(q, path=1)
        assert np.abs(gc - c[None, :]).max() <= 2e-6 * np.abs(c).max() + 1e-6
    finally:
        sc.close()


# --------------------------------------------------------------------------------- full step
def test_decode_step_device_and_host_entry():
    """append + decode of a real step (GPT-2 124M shape, 2 layers), device-resident and through
    the host-buffer entry pa_decode_step_host; K/V written by the append kernel are what the
    decode kernel then reads."""
    NH, hs, bs, B = 12, 64, 16, 9
    Cc = NH * hs
    ctx0 = [0, 1, 15, 16, 31, 32, 100, 255, 256]
    sc = Scenario(NH, hs, bs, ctx0, n_layers=2, layer=1, seed=51, shuffle=False, extra_blocks=32)
    try:
        eng, orc = sc.eng, sc.orc
        for step in range(3):
            qkv = oa.normal((B, 3 * Cc), seed=200 + step)
            assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
            for s in range(B):
                orc.add_to_cache(qkv[s][None, None, :], 1, 1, 1, prompt=s)
            want = orc.decode_batch(sc.seq_ids, NH, qkv[:, :Cc])
            if step % 2 == 0:      # device-resident API
                pa.check(eng.upload(), "upload")
                d = pa.DevBuf.from_numpy(qkv)
                o = pa.DevBuf(B * Cc * 4)
                pa.check(eng.append(1, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc), "append")
                pa.check(eng.decode(1, d.ptr, 3 * Cc, o.ptr, Cc), "decode")
                eng.sync()
                got = o.download((B, Cc))
            elif step == 1:        # pageable host buffers in, host buffers out (staged copies)
                got = np.zeros((B, Cc), dtype=np.float32)
                pa.check(eng.decode_step_host(1, qkv.ctypes.data, got.ctypes.data), "decode_step_host")
            assert_close(got, want, f"step {step}")
        # pinned host buffers: the kernel reads/writes them directly (zero-copy), and the staged
        # variant of the same call
        import ctypes as Ct
        lib = eng.lib
        hin, hout = lib.pa_host_alloc(B * 3 * Cc * 4), lib.pa_host_alloc(B * Cc * 4)
        try:
            for step, nozc in ((10, 0), (11, 1)):
                qkv = oa.normal((B, 3 * Cc), seed=200 + step)
                Ct.memmove(hin, qkv.ctypes.data, qkv.nbytes)
                eng.tune(pa.PA_TUNE_NO_ZEROCOPY, nozc)
                assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
                for s in range(B):
                    orc.add_to_cache(qkv[s][None, None, :], 1, 1, 1, prompt=s)
                want = orc.decode_batch(sc.seq_ids, NH, qkv[:, :Cc])
                pa.check(eng.decode_step_host(1, hin, hout), "decode_step_host pinned")
                got = np.ctypeslib.as_array(Ct.cast(hout, Ct.POINTER(Ct.c_float)), (B, Cc)).copy()
                assert_close(got, want, f"pinned host entry, zero-copy={'off' if nozc else 'on'}")
                k, v = eng.read_pool_rows(1, eng.slot_mapping())
                assert np.array_equal(k, qkv[:, Cc:2 * Cc]) and np.array_equal(v, qkv[:, 2 * Cc:])
        finally:
            lib.pa_host_free(hin); lib.pa_host_free(hout)
    finally:
        sc.close()


def test_decode_step_host_async_matches_sync():
    """pa_decode_step_host_async: layers queued on the handle's stream, one sync; same results as the
    synchronous entry (zero-copy and staged), pageable buffers are refused."""
    import ctypes as Ct
    NH, hs, bs, B, L = 4, 64, 16, 6, 3
    Cc = NH * hs
    sc = Scenario(NH, hs, bs, [33, 1, 16, 70, 5, 48], n_layers=L, layer=0, seed=77, extra_blocks=16)
    lib = sc.eng.lib
    hin = [lib.pa_host_alloc(B * 3 * Cc * 4) for _ in range(L)]
    hout = [lib.pa_host_alloc(B * Cc * 4) for _ in range(2 * L)]
    try:
        eng = sc.eng
        for l in range(L):
            arr = np.ctypeslib.as_array(Ct.cast(hin[l], Ct.POINTER(Ct.c_float)), (B, 3 * Cc))
            arr[:] = oa.normal((B, 3 * Cc), seed=80 + l)
        results = {}
        for mode in ("sync", "async", "async-staged"):
            eng.tune(pa.PA_TUNE_NO_ZEROCOPY, 1 if mode == "async-staged" else 0)
            assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
            outs = hout[:L] if mode == "sync" else hout[L:]
            for l in range(L):
                fn = eng.decode_step_host if mode == "sync" else eng.decode_step_host_async
                pa.check(fn(l, hin[l], outs[l]), mode)
            pa.check(lib.pa_decode_step_host_sync(eng.h), "sync")
            results[mode] = [np.ctypeslib.as_array(Ct.cast(o, Ct.POINTER(Ct.c_float)), (B, Cc)).copy() for o in outs]
            pa.check(eng.step_rollback(), "rollback")
        for l in range(L):
            assert np.array_equal(results["sync"][l], results["async"][l])
            assert np.array_equal(results["sync"][l], results["async-staged"][l])
        pageable = np.zeros((B, 3 * Cc), dtype=np.float32)
        assert eng.step_begin(sc.seq_ids, [1] * B) == 0
        assert eng.decode_step_host_async(0, pageable.ctypes.data, hout[0]) == pa.PA_ERR_INVALID
    finally:
        for p in hin + hout:
            lib.pa_host_free(p)
        sc.close()


def test_decode_step_host_layers_pipelined_with_tickets():
    """pa_decode_step_host_layers_async (every layer of a step in one call) in its three modes -- whole-step staged
    copies (auto for a small step / forced), zero-copy -- with the host running one step AHEAD: step n+1 is queued
    (new tables, new layers) before step n's ticket is waited for.  Every step's output rows equal what the
    device-resident pa_decode_append gives for the same step; strided per-layer rows; stale tickets are refused."""
    import ctypes as Ct
    NH, hs, bs, B, L = 4, 64, 16, 6, 3
    Cc = NH * hs
    n_steps = 5
    lib = pa.load()
    in_stride, out_stride = B * 3 * Cc + 64, B * Cc + 32          # padded layer strides: the 2-D copy path
    hin = lib.pa_host_alloc(n_steps * L * in_stride * 4)
    hout = lib.pa_host_alloc(n_steps * L * out_stride * 4)
    ins = np.ctypeslib.as_array(Ct.cast(hin, Ct.POINTER(Ct.c_float)), (n_steps, L, in_stride))
    outs = np.ctypeslib.as_array(Ct.cast(hout, Ct.POINTER(Ct.c_float)), (n_steps, L, out_stride))
    ins[:] = oa.normal((n_steps, L, in_stride), seed=123)
    try:
        want = None
        for mode in ("device", 0, 3, 2, 1):
            sc = Scenario(NH, hs, bs, [33, 1, 16, 70, 5, 48], n_layers=L, layer=0, seed=77, extra_blocks=32)
            eng = sc.eng
            try:
                got = np.zeros((n_steps, L, B, Cc), dtype=np.float32)
                if mode == "device":
                    for st in range(n_steps):
                        assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
                        pa.check(eng.upload(), "upload")
                        for l in range(L):
                            d = pa.DevBuf.from_numpy(np.ascontiguousarray(ins[st, l, :B * 3 * Cc].reshape(B, 3 * Cc)))
                            o = pa.DevBuf(B * Cc * 4)
                            pa.check(eng.decode_append(l, d.ptr, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc, o.ptr, Cc), "decode_append")
                            eng.sync()
                            got[st, l] = o.download((B, Cc))
                            d.free(); o.free()
                    want = got
                    continue
                eng.tune(pa.PA_TUNE_NO_ZEROCOPY, mode)
                outs[:] = np.nan
                tickets = []
                for st in range(n_steps):
                    assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
                    pa.check(lib.pa_decode_step_host_layers_async(eng.h, ins[st].ctypes.data, in_stride, outs[st].ctypes.data, out_stride),
                             f"layers_async mode {mode}")
                    t = lib.pa_decode_step_host_mark(eng.h)
                    assert t > 0
                    tickets.append(t)
                    if st >= 1:                      # one step behind: wait for step st-1 while step st runs
                        pa.check(lib.pa_decode_step_host_wait(eng.h, tickets[st - 1]), "wait")
                        got[st - 1] = outs[st - 1, :, :B * Cc].reshape(L, B, Cc)
                pa.check(lib.pa_decode_step_host_wait(eng.h, tickets[-1]), "wait")
                got[-1] = outs[-1, :, :B * Cc].reshape(L, B, Cc)
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"host mode {mode}"
                assert np.isnan(outs[:, :, B * Cc:]).all(), "the padding between the layers' output rows was written"
                assert lib.pa_decode_step_host_wait(eng.h, tickets[-1] + 1) == pa.PA_ERR_INVALID      # never made
                for _ in range(9):
                    lib.pa_decode_step_host_mark(eng.h)
                assert lib.pa_decode_step_host_wait(eng.h, tickets[0]) == pa.PA_ERR_INVALID            # too old
                pa.check(lib.pa_decode_step_host_sync(eng.h), "sync")
            finally:
                sc.close()
    finally:
        lib.pa_host_free(hin)
        lib.pa_host_free(hout)


# --------------------------------------------------------------------------------- prefill rows
def _run_prefill(NH, hs, bs, before, n_new, path, seed=61, kv_start=None, dist="normal", shuffle=False, nwg=0, bn=0):
    """Append + causal rows for a mixed batch; returns (got, want32, scenario-free copies)."""
    Cc = NH * hs
    sc = Scenario(NH, hs, bs, before, seed=seed, extra_blocks=sum((n + bs - 1) // bs + 1 for n in n_new) + 8,
                  max_batch_tokens=sum(n_new), dist=dist, shuffle=shuffle)   # no eviction
    try:
        eng, orc = sc.eng, sc.orc
        ntok = sum(n_new)
        if dist == "normal":
            qkv = oa.normal((ntok, 3 * Cc), seed=seed + 1)
        else:
            qkv = oa.uniform((ntok, 3 * Cc), 0.0, 100.0, seed=seed + 1)
        eng.tune(pa.PA_TUNE_PREFILL_PATH, path)
        eng.tune(pa.PA_TUNE_TC_WARPGROUPS, nwg)
        eng.tune(pa.PA_TUNE_TC_KEY_TILE, bn)
        assert eng.step_begin(sc.seq_ids, n_new) == 0, pa.last_error()
        if kv_start is not None:
            assert eng.step_set_kv_start(kv_start) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        d = pa.DevBuf.from_numpy(qkv)
        o = pa.DevBuf(ntok * Cc * 4)
        pa.check(eng.lib.pa_memset(o.ptr, 0xff, ntok * Cc * 4, None), "memset")   # NaN canary
        pa.check(eng.append(0, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc), "append")
        pa.check(eng.prefill(0, d.ptr, 3 * Cc, o.ptr, Cc), "prefill")
        eng.sync()
        got = o.download((ntok, Cc))
        row = 0
        for s, n in enumerate(n_new):
            for _ in range(n):
                orc.add_to_cache(qkv[row][None, None, :], 1, 1, 1, prompt=s)
                row += 1
        ks = [0] * len(n_new) if kv_start is None else kv_start
        want = orc.attend_rows(sc.seq_ids, ks, [b + 1 for b in before], n_new, NH, qkv[:, :Cc])
        return got, want
    finally:
        sc.close()


@pytest.mark.parametrize("path", [1, 2], ids=["tiled", "rows"])
@pytest.mark.parametrize("NH,hs,bs", [(12, 64, 16), (2, 5, 2), (4, 128, 32)])
def test_prefill_rows_match_oracle(NH, hs, bs, path):
    """Causal rows through the block table: prompt prefill (ctx_before = 0) and chunked prefill
    on top of cached tokens, mixed in one batch."""
    if path == 1 and hs not in (64, 128):
        pytest.skip("tiled kernel: head_dim 64/128")
    got, want = _run_prefill(NH, hs, bs, [0, 5, 40, 0], [33, 7, 1, 64], path)
    assert_close(got, want, "prefill rows")


@pytest.mark.parametrize("NH,hs,bs,before,n_new", [
    (3, 64, 16, [0, 100, 0, 17, 300], [300, 129, 128, 1, 257]),     # several q tiles, diagonal + tail tiles
    (2, 128, 16, [0, 77, 0], [200, 65, 64]),
    (2, 64, 4, [3, 0], [130, 70]),                                   # small pages
    (2, 64, 3, [5, 0], [100, 64]),                                   # block size not a power of two
    (1, 64, 32, [0], [1]),                                           # a single row
])
def test_prefill_tiled_shapes(NH, hs, bs, before, n_new):
    got, want = _run_prefill(NH, hs, bs, before, n_new, 1, shuffle=True)
    assert_close(got, want, "tiled prefill")


@pytest.mark.parametrize("nwg,bn", [(1, 0), (2, 0), (1, 128), (2, 128)])
@pytest.mark.parametrize("NH,hs,bs,before,n_new", [
    (3, 64, 16, [0, 100, 0, 17, 300], [300, 129, 128, 1, 257]),
    (2, 128, 16, [0, 77, 0], [200, 65, 64]),
    (2, 64, 8, [3, 0], [130, 70]),
    (2, 64, 32, [0, 500], [1000, 3]),
    (1, 128, 64, [0], [129]),
])
def test_prefill_tcgen05_tf32(NH, hs, bs, before, n_new, nwg, bn):
    """Opt-in tensor-core prefill (tcgen05 kind::tf32, TMEM accumulators, TMA page gather) against
    the fp32 oracle at the TF32 tolerance stated in gpu_common.TC_REL_TOL."""
    if bn == 128 and hs != 64:
        pytest.skip("key-tile knob applies to head_dim 64")
    got, want = _run_prefill(NH, hs, bs, before, n_new, 3, shuffle=True, nwg=nwg, bn=bn)
    err = assert_close_tc(got, want, "tcgen05 prefill")
    print(f"tcgen05 tf32 prefill hs={hs} bs={bs} nwg={nwg} bn={bn}: max rel err {err:.2e}")


@pytest.mark.parametrize("NH,hs,bs,before,n_new", [
    (3, 64, 16, [0, 100, 0, 17, 300], [300, 129, 128, 1, 257]),     # several q tiles, diagonal + tail tiles, a 1-row sequence
    (2, 128, 16, [0, 77, 0], [200, 65, 64]),
    (2, 64, 8, [3, 0], [130, 70]),
    (2, 64, 32, [0, 500], [1000, 3]),
    (2, 64, 64, [0, 130], [260, 33]),                                # a page = a whole key tile
    (1, 128, 32, [0], [129]),
    (1, 128, 8, [1000], [700]),                                      # chunk on top of a long cache, head_dim 128
    (12, 64, 16, [0], [2048]),                                       # GPT-2 124M heads, a 2048-token prompt
])
def test_prefill_tcgen05_3xtf32_is_fp32_accurate(NH, hs, bs, before, n_new):
    """The fp32-ACCURATE tensor-core prefill (pa_prefill_tc3.cu: tcgen05 kind::tf32 with the 3xTF32 split, fresh
    TMEM accumulators per key tile, register accumulation across tiles, expf) against the fp32 oracle at the
    path's tolerance: max|a-b|/max|ref| <= 1e-5 and allclose(rtol 1e-5, atol 3e-6) (gpu_common.assert_close_tc3)."""
    got, want = _run_prefill(NH, hs, bs, before, n_new, 4, shuffle=True)
    err = assert_close_tc3(got, want, "tcgen05 3xTF32 prefill")
    print(f"tcgen05 3xtf32 prefill NH={NH} hs={hs} bs={bs}: max rel err {err:.2e}")


def test_prefill_tcgen05_3xtf32_persistent_schedule():
    """The kernel is persistent: one CTA per SM walks a host-built list of (sequence, q tile, head) units with nothing
    draining in between (the next Q tile prefetched through the staging tile, the previous unit's rows stored after the
    next unit's first key tile).  The cases above mostly give a CTA one unit; these give every CTA several: 40 ragged
    sequences x 6 heads = 480+ units on 148 SMs, cached context and fragmented pages included; and a sliding window
    that leaves whole q tiles without a visible key (zero rows, no key tile: every role must skip them alike)."""
    rng = np.random.default_rng(5)
    n_new = [int(x) for x in rng.integers(1, 300, 38)] + [700, 513]
    before = [int(x) for x in rng.integers(0, 200, 38)] + [0, 90]
    got, want = _run_prefill(6, 64, 16, before, n_new, 4, shuffle=True)
    assert_close_tc3(got, want, "3xTF32 prefill, several units per CTA")
    got, want = _run_prefill(2, 128, 16, [0, 40, 0] * 30, [200, 129, 260] * 30, 4, shuffle=True)
    assert_close_tc3(got, want, "3xTF32 prefill, several units per CTA, head_dim 128")
    # window start beyond the first q tile's last row: its rows, and rows 128..199 of the second tile, see nothing
    got, want = _run_prefill(3, 64, 16, [0, 0], [300, 450], 4, kv_start=[200, 129])
    assert (got[:200] == 0).all() and (got[300:300 + 129] == 0).all()
    assert_close_tc3(got, want, "3xTF32 prefill, q tiles without keys")


def _prefill_two_paths(NH, hs, bs, before, n_new, kv_start, seed, paths=(4, 2)):
    """The same step through two kernels on ONE engine (the step is rolled back in between): outputs of both."""
    Cc = NH * hs
    sc = Scenario(NH, hs, bs, before, seed=seed, extra_blocks=sum((n + bs - 1) // bs + 1 for n in n_new) + 8,
                  max_batch_tokens=sum(n_new), shuffle=True)
    try:
        eng = sc.eng
        ntok = sum(n_new)
        qkv = oa.normal((ntok, 3 * Cc), seed=seed + 1)
        d = pa.DevBuf.from_numpy(qkv)
        o = pa.DevBuf(ntok * Cc * 4)
        outs = []
        for path in paths:
            eng.tune(pa.PA_TUNE_PREFILL_PATH, path)
            assert eng.step_begin(sc.seq_ids, n_new) == 0, pa.last_error()
            if kv_start is not None:
                assert eng.step_set_kv_start(kv_start) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            pa.check(eng.lib.pa_memset(o.ptr, 0xff, ntok * Cc * 4, None), "memset")   # NaN canary
            pa.check(eng.append(0, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc), "append")
            pa.check(eng.prefill(0, d.ptr, 3 * Cc, o.ptr, Cc), "prefill")
            eng.sync()
            outs.append(o.download((ntok, Cc)))
            pa.check(eng.step_rollback(), "rollback")
        d.free(); o.free()
        return outs
    finally:
        sc.close()


@pytest.mark.timeout(300)
def test_prefill_tcgen05_3xtf32_random_ragged_steps_match_rows_kernel():
    """Fuzz of the persistent kernel's scheduling (units per CTA, unit boundaries, q tiles without keys, one-tile
    units, partial q tiles, both head dims, every page size of its domain, sliding windows) against the generic rows
    kernel, which the cases above pin to the oracle: 30 seeded random steps, same tolerance as against the oracle."""
    rng = np.random.default_rng(int(os.environ.get("PA_FUZZ_SEED", "2024")))
    for case in range(int(os.environ.get("PA_FUZZ_CASES", "30"))):
        hs = int(rng.choice([64, 128]))
        bs = int(rng.choice([8, 16, 32, 64] if hs == 64 else [8, 16, 32]))
        NH = int(rng.integers(1, 7))
        B = int(rng.integers(1, 25))
        shape = rng.integers(0, 3)
        if shape == 0:      # many short sequences
            n_new = [int(x) for x in rng.integers(1, 140, B)]
            before = [int(x) for x in rng.integers(0, 100, B)]
        elif shape == 1:    # few rows on longer caches (the verify step of speculative decoding)
            n_new = [int(x) for x in rng.integers(1, 9, B)]
            before = [int(x) for x in rng.integers(0, 900, B)]
        else:               # a few long prompts
            B = min(B, 4)
            n_new = [int(x) for x in rng.integers(100, 700, B)]
            before = [int(x) for x in rng.integers(0, 300, B)]
        kv_start = None
        if rng.random() < 0.4:        # windows; some start beyond what the first rows may see (rows without keys)
            kv_start = [int(rng.integers(0, before[i] + n_new[i])) for i in range(B)]
        got, want = _prefill_two_paths(NH, hs, bs, before, n_new, kv_start, seed=100 + case)
        assert np.isfinite(want).all()
        assert_close_tc3(got, want, f"case {case}: NH={NH} hs={hs} bs={bs} B={B} n_new={n_new[:6]} before={before[:6]} kv_start={None if kv_start is None else kv_start[:6]}")


@pytest.mark.timeout(300)
@pytest.mark.parametrize("name,NH,hs,B,T,before", [
    ("headline-16x2048", 12, 64, 16, 2048, 0),                 # bench.py's prefill shape
    ("cfg5-chunk-on-30000-cached", 32, 128, 2, 2048, 30000),   # a 2048-token chunk of the 32k prompts, head_dim 128
])
def test_prefill_full_size_properties(name, NH, hs, B, T, before):
    """The persistent 3xTF32 prefill at the sizes bench.py times, through properties that need no CPU pass over 10^11
    products: (1) with every V row equal to one vector c, every output row is c (the weights sum to one; all rows, all
    heads: masking and normalisation); (2) the last row of every sequence equals what the DECODE kernel computes for
    that query over the same cache (two independent kernels, one of them pinned to the oracle at this size by
    test_baseline_full_size_decode_parity_and_properties); (3) causality, bit-exact: other K/V in the last 100 tokens of
    every sequence leaves all earlier rows' outputs unchanged; (4) the run is deterministic."""
    lib = pa.load()
    bs = 16
    Cc = NH * hs
    pages = (before + T + bs - 1) // bs + 1
    eng = pa.PagedAttn(bs, B * pages + 8, B, NH, hs, n_layers=1, device=0, max_batch_tokens=B * T)
    try:
        stream = lib.pa_stream_of(eng.h)
        rng = np.random.default_rng(11)
        perm = rng.permutation(B * pages + 8)
        nb = (before + bs - 1) // bs
        pool_floats = (B * pages + 8) * bs * Cc
        pa.check(lib.pa_fill_normal(eng.pool_k(0), pool_floats, 1.0, 0.0, 501, stream), "fill")
        pa.check(lib.pa_fill_normal(eng.pool_v(0), pool_floats, 1.0, 0.0, 502, stream), "fill")
        for s_ in range(B):
            if before:
                assert eng.seq_adopt(s_, perm[s_ * pages: s_ * pages + nb], before) == 0, pa.last_error()
        ntok = B * T
        d_in, d_o = pa.DevBuf(ntok * 3 * Cc * 4), pa.DevBuf(ntok * Cc * 4)
        pa.check(lib.pa_fill_normal(d_in.ptr, ntok * 3 * Cc, 1.0, 0.0, 503, stream), "fill")
        seq_ids = np.arange(B, dtype=np.int32)
        eng.tune(pa.PA_TUNE_PREFILL_PATH, 4)

        def run_step(keep=False):
            assert eng.step_begin(seq_ids, [T] * B) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            pa.check(lib.pa_memset(d_o.ptr, 0xff, ntok * Cc * 4, None), "memset")
            pa.check(eng.append(0, d_in.ptr + Cc * 4, d_in.ptr + 2 * Cc * 4, 3 * Cc), "append")
            pa.check(eng.prefill(0, d_in.ptr, 3 * Cc, d_o.ptr, Cc), "prefill")
            eng.sync()
            out = d_o.download((ntok, Cc))
            if not keep:
                pa.check(eng.step_rollback(), "rollback")
            return out

        base = run_step()
        assert np.isfinite(base).all()
        assert np.array_equal(base, run_step()), "not deterministic"                                  # (4)

        # (3) other K/V in the last 100 tokens of every sequence: rows before them unchanged, bit for bit
        tail = oa.normal((100, 3 * Cc), seed=77)
        tail[:, :Cc] = 0.0
        saved = []
        for s_ in range(B):
            off = ((s_ * T + T - 100) * 3 * Cc) * 4
            row = d_in.download((100, 3 * Cc), offset_bytes=off)
            saved.append(row)
            mod = row.copy()
            mod[:, Cc:] = tail[:, Cc:]                           # K and V columns only: the queries stay
            pa.check(lib.pa_memcpy_h2d(d_in.ptr + off, mod.ctypes.data, mod.nbytes, None), "h2d")
        changed = run_step()
        for s_ in range(B):
            assert np.array_equal(changed[s_ * T: s_ * T + T - 100], base[s_ * T: s_ * T + T - 100]), f"sequence {s_}: a future token changed a past row"
            assert not np.array_equal(changed[s_ * T + T - 100: (s_ + 1) * T], base[s_ * T + T - 100: (s_ + 1) * T])
            pa.check(lib.pa_memcpy_h2d(d_in.ptr + ((s_ * T + T - 100) * 3 * Cc) * 4, saved[s_].ctypes.data, saved[s_].nbytes, None), "h2d")

        # (2) last rows against the decode kernel over the same cache (the step stays appended)
        full = run_step(keep=True)
        assert np.array_equal(full, base)
        q_last = np.stack([d_in.download((1, 3 * Cc), offset_bytes=((s_ * T + T - 1) * 3 * Cc) * 4)[0, :Cc] for s_ in range(B)])
        d_q, d_dec = pa.DevBuf.from_numpy(np.ascontiguousarray(q_last)), pa.DevBuf(B * Cc * 4)
        assert eng.step_begin_readonly(seq_ids) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        pa.check(eng.decode(0, d_q.ptr, Cc, d_dec.ptr, Cc), "decode")
        eng.sync()
        dec = d_dec.download((B, Cc))
        err = assert_close_tc3(full[T - 1::T], dec, f"{name}: last prefill rows vs the decode kernel")
        print(f"{name}: last prefill rows vs decode kernel, max rel err {err:.2e}")
        d_q.free(); d_dec.free()
        for s_ in range(B):
            eng.seq_free(s_)

        # (1) constant V rows (cache and new tokens): every output row is that vector
        c = oa.normal((1, Cc), seed=5)[0]
        vrows = np.ascontiguousarray(np.broadcast_to(c, (4096, Cc)))
        for r0 in range(0, (B * pages + 8) * bs, 4096):
            n = min(4096, (B * pages + 8) * bs - r0)
            pa.check(lib.pa_memcpy_h2d(eng.pool_v(0) + r0 * Cc * 4, vrows.ctypes.data, n * Cc * 4, None), "h2d")
        chunk = d_in.download((T, 3 * Cc), offset_bytes=0)
        for s_ in range(B):
            if before:
                assert eng.seq_adopt(s_, perm[s_ * pages: s_ * pages + nb], before) == 0, pa.last_error()
            blk = d_in.download((T, 3 * Cc), offset_bytes=(s_ * T * 3 * Cc) * 4) if s_ else chunk
            blk[:, 2 * Cc:] = c
            pa.check(lib.pa_memcpy_h2d(d_in.ptr + (s_ * T * 3 * Cc) * 4, blk.ctypes.data, blk.nbytes, None), "h2d")
        const = run_step()
        scale = np.abs(c).max()
        dev = np.abs(const.astype(np.float64) - c.astype(np.float64)).max() / scale
        assert dev <= 1e-5, f"{name}: constant-V rows deviate by {dev:.2e} of max|c|"
        d_in.free(); d_o.free()
    finally:
        eng.close()


def test_prefill_tcgen05_3xtf32_window_large_logits_and_domain():
    got, want = _run_prefill(2, 64, 16, [150, 70], [90, 140], 4, kv_start=[37, 64])
    assert_close_tc3(got, want, "3xTF32 prefill, window")
    # the reference test's U[0,100) inputs (logits ~1e5): as well conditioned as the fp32 kernels are (see gpu_common)
    got_t, want = _run_prefill(2, 64, 16, [0, 20], [70, 40], 4, dist="uniform")
    got_r, _ = _run_prefill(2, 64, 16, [0, 20], [70, 40], 2, dist="uniform")
    assert np.isfinite(got_t).all()
    ref_gap = np.abs(got_r.astype(np.float64) - want).max()
    assert np.abs(got_t.astype(np.float64) - want).max() <= max(4 * ref_gap, 1e-2)
    with pytest.raises(pa.PagedAttnError):          # block size 4 < one swizzle group: forced path fails loudly
        _run_prefill(2, 64, 4, [0], [40], 4)
    with pytest.raises(pa.PagedAttnError):          # head_dim 128: pages above 32 tokens exceed the 32-key tile
        _run_prefill(1, 128, 64, [0], [129], 4)
    # the automatic choice (path 0) is this kernel wherever its domain allows, small steps included, and the SIMT
    # kernels outside it (pages of 4 tokens; head_dim 5): same answers
    big, want_big = _run_prefill(2, 64, 16, [0, 10], [300, 200], 0)
    assert_close_tc3(big, want_big, "auto, large step")
    small, want_small = _run_prefill(2, 64, 16, [0, 10], [20, 9], 0)
    assert_close_tc3(small, want_small, "auto, small step")
    few, want_few = _run_prefill(12, 64, 16, [1000] * 20, [4] * 20, 0, shuffle=True)     # few rows on a long cache
    assert_close_tc3(few, want_few, "auto, 4 rows on 1000 cached")
    tiny_pages, want_tp = _run_prefill(2, 64, 4, [0, 10], [40, 9], 0)
    assert_close(tiny_pages, want_tp, "auto, pages of 4 tokens (SIMT)")
    odd, want_odd = _run_prefill(2, 5, 2, [3], [17], 0)
    assert_close(odd, want_odd, "auto, head_dim 5 (rows kernel)")


def test_prefill_tcgen05_window_and_unsupported_shapes():
    got, want = _run_prefill(2, 64, 16, [150, 70], [90, 140], 3, kv_start=[37, 64])
    assert_close_tc(got, want, "tcgen05 prefill, window")
    with pytest.raises(pa.PagedAttnError):          # block size 4 < one swizzle group: fails loudly, no silent fallback
        _run_prefill(2, 64, 4, [0], [40], 3)


def test_prefill_tiled_sliding_window():
    """kv_start > 0 (the reference's `offset`): every row of the chunk sees [kv_start, its own token]."""
    got, want = _run_prefill(2, 64, 16, [150, 70], [90, 140], 1, kv_start=[37, 64])
    assert_close(got, want, "tiled prefill, window")


def test_prefill_tiled_equals_rows_kernel_large_logits():
    """Reference test value range U[0,100): the two kernels must agree with each other about as
    well as either agrees with the fp32 reference (ill-conditioned softmax, see gpu_common)."""
    got_t, want = _run_prefill(2, 64, 16, [0, 20], [70, 40], 1, dist="uniform")
    got_r, _ = _run_prefill(2, 64, 16, [0, 20], [70, 40], 2, dist="uniform")
    assert np.isfinite(got_t).all()
    ref_gap = np.abs(got_r.astype(np.float64) - want).max()
    assert np.abs(got_t.astype(np.float64) - want).max() <= max(4 * ref_gap, 1e-2)


# --------------------------------------------------------------------------------- compat API
def _compat_manager(lib, C_, bs, mb, mp):
    lib.pa_set_default_geometry(bs, mb, mp)
    m = lib.create_block_manager(C_)
    assert m, "create_block_manager failed"
    return m


@pytest.mark.parametrize("device_buffers", [False, True], ids=["host-buffers", "device-buffers"])
def test_compat_sliding_window_like_paged_infer_main(device_buffers):
    """The reference call site (paged_infer.c:710-715) and main's pattern (:1055-1057): T=32,
    first add_to_cache(n_tail=T), then n_tail=1 with offset=1..18, crossing into a second page.
    Drop-in names, host buffers as in the reference."""
    lib = pa.load()
    bs, mb, mp, T, Cc, NH = 32, 100, 100, 32, 768, 12
    m = _compat_manager(lib, Cc, bs, mb, mp)
    orc = oa.OrcManager(Cc, bs, mb, mp)
    ref = oa.RefManager(Cc, bs, mb, mp) if oa.have_ref(bs, mb, mp) else None
    try:
        stream = oa.normal((T + 18, 3 * Cc), seed=2024)
        for step in range(19):
            window = np.ascontiguousarray(stream[step:step + T][None])
            n_tail = T if step == 0 else 1
            out = np.zeros((1, T, Cc), dtype=np.float32)
            if device_buffers:
                d_in = pa.DevBuf.from_numpy(window)
                d_out = pa.DevBuf(out.nbytes)
                lib.add_to_cache(m, d_in.ptr, 1, T, Cc, n_tail)
            else:
                lib.add_to_cache(m, window.ctypes.data, 1, T, Cc, n_tail)
            nb = C.c_int()
            kv = lib.collect_kv_blocks(m, 0, C.byref(nb))
            assert kv
            if device_buffers:
                lib.attention_paged(d_out.ptr, None, None, d_in.ptr, kv[0], kv[1], 1, T, Cc, NH, step)
                out = d_out.download(out.shape)
            else:
                lib.attention_paged(out.ctypes.data, None, None, window.ctypes.data, kv[0], kv[1], 1, T, Cc, NH, step)
            orc.add_to_cache(window, 1, T, n_tail)
            _, want = orc.attend(0, window, 1, T, NH, step)
            assert nb.value == len(orc.table(0))
            assert_close(out, want, f"compat step {step}")
            if ref is not None:
                ref.add_to_cache(window, 1, T, n_tail)
                _, rwant = ref.attend(0, window, 1, T, NH, step)
                assert_close(out, rwant, f"compat vs compiled reference, step {step}")
                assert ref.table(0) == pa.ManagerAdapter(m).table(0)
                assert ref.epoch() == m.contents.lru_epoch
        ad = pa.ManagerAdapter(m)
        assert ad.table(0) == [0, 1]
        assert [ad.block_info(i)[0] for i in (0, 1)] == [32, 18]
        assert m.contents.lru_epoch == 19
    finally:
        lib.destroy_block_manager(m)
        orc.close()
        if ref is not None:
            ref.close()
        lib.pa_set_default_geometry(32, 100, 100)


def test_compat_reference_unit_test_shape():
    """test_paged_attn.c:184-188 (B=1,T=20,C=10,NH=2, 10 blocks of 2) on U[0,100) inputs, through
    the compat names; block_manager_test.c's write/readback pattern on device pages."""
    lib = pa.load()
    bs, mb, mp, T, Cc, NH = 2, 64, 8, 20, 10, 2
    m = _compat_manager(lib, Cc, bs, mb, mp)
    orc = oa.OrcManager(Cc, bs, mb, mp)
    try:
        inp = oa.uniform((1, T, 3 * Cc), 0.0, 100.0, seed=42)
        for t0 in range(0, T, bs):
            blk = lib.request_block(m, 0)
            assert blk
            k_rows = np.ascontiguousarray(inp[0, t0:t0 + bs, Cc:2 * Cc])
            v_rows = np.ascontiguousarray(inp[0, t0:t0 + bs, 2 * Cc:])
            pa.check(lib.pa_memcpy_h2d(blk.contents.keys, k_rows.ctypes.data, k_rows.nbytes, None), "h2d")
            pa.check(lib.pa_memcpy_h2d(blk.contents.values, v_rows.ctypes.data, v_rows.nbytes, None), "h2d")
            blk.contents.filled = bs
            back = np.zeros_like(k_rows)
            pa.check(lib.pa_memcpy_d2h(back.ctypes.data, blk.contents.keys, back.nbytes, None), "d2h")
            assert np.array_equal(back, k_rows)                      # block_manager_test.c:31-38
            oidx = orc.request_block(0)
            ok, ov = orc.page_arrays(oidx)
            ok[:], ov[:] = k_rows, v_rows
            orc.set_filled(oidx, bs)
        nb = C.c_int()
        kv = lib.collect_kv_blocks(m, 0, C.byref(nb))
        assert nb.value == 10
        out = np.zeros((1, T, Cc), dtype=np.float32)
        lib.attention_paged(out.ctypes.data, None, None, inp.ctypes.data, kv[0], kv[1], 1, T, Cc, NH, 0)
        _, want = orc.attend(0, inp, 1, T, NH, 0)
        assert np.abs(out - want).max() <= 1e-2                      # the reference's own tolerance
        assert_close(out, want, "compat test_paged_attn shape")
        # error behaviour of the reference API
        assert not lib.request_block(m, 100) and not lib.request_block(m, -1)
        assert not lib.collect_kv_blocks(m, 5, C.byref(nb)) and nb.value == 0
        lib.free_blocks_for_prompt(m, 0)
        assert m.contents.prompt_block_count[0] == 0
    finally:
        lib.destroy_block_manager(m)
        orc.close()
        lib.pa_set_default_geometry(32, 100, 100)
