"""-m gpu parity tests of the row next to the path: the QKV projection with the KV
append fused into its epilogue (pa_qkv_append), and the reference-named matmul_forward /
matmul_cached (paged_infer.c:92-160).  Checker: oracle/paged_oracle.c (orc_matmul_*, pinned
bit-exact to the compiled reference by tests/test_oracle_pinned.py)."""
import ctypes as C

import numpy as np
import pytest

import oracle_api as oa
from gpu_common import REL_TOL, Scenario, assert_close, pa

pytestmark = pytest.mark.gpu


def assert_close_gemm(got, want, what=""):
    """GEMM outputs: max|a-b| / max|ref| <= 1e-5 (north_star's bar).  The element-wise allclose of
    gpu_common.assert_close (atol 1e-6) does not apply: a K-term fp32 dot product is itself only
    defined to ~sqrt(K)*6e-8*|terms| (~1e-6 here) under a change of summation order, which both the
    -Ofast reference and any tiled kernel make."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape and np.isfinite(got).all(), what
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    assert err <= REL_TOL, f"{what}: max|a-b|/max|ref| = {err:.3e} > {REL_TOL}"
    return err


def _oracle_matmul(x, w, bias, cached=False, T=1):
    ol = oa.load_oracle()
    rows, Cc = x.shape
    OC = w.shape[0]
    out = np.zeros((rows, OC), dtype=np.float32)
    fn = ol.orc_matmul_cached if cached else ol.orc_matmul_forward
    fn(oa.fptr(out), oa.fptr(x), oa.fptr(w), oa.fptr(bias) if bias is not None else None, rows // T, T, Cc, OC)
    return out


@pytest.mark.parametrize("path", [1, 2], ids=["simt", "tcgen05-3xtf32"])
@pytest.mark.parametrize("NH,hs,bs,ctx", [
    (12, 64, 16, [1, 16, 17, 100, 333, 64, 5]),
    (25, 64, 16, [40, 1, 129]),
    (4, 128, 32, [31, 32, 33, 200]),
    (3, 20, 8, [9, 2]),                        # C = 60: not a multiple of the tile sizes
])
def test_qkv_append_then_decode_matches_oracle(NH, hs, bs, ctx, path):
    """One decode step: x -> (q | k | v) with k, v written straight to the page slots by the GEMM
    epilogue, then paged decode attention.  Oracle: matmul_forward (the single-row case of
    matmul_cached) -> add_to_cache -> attention row."""
    Cc = NH * hs
    B = len(ctx)
    before = [c - 1 for c in ctx]
    if path == 2 and (NH * hs) % 32:
        pytest.skip("tcgen05 GEMM: C % 32 == 0")
    sc = Scenario(NH, hs, bs, before, seed=71, extra_blocks=B + 8)
    try:
        eng, orc, lib = sc.eng, sc.orc, sc.eng.lib
        eng.tune(pa.PA_TUNE_GEMM_PATH, path)
        x = oa.normal((B, Cc), seed=72)
        w = (oa.normal((3 * Cc, Cc), seed=73) * np.float32(1.0 / np.sqrt(Cc))).astype(np.float32)
        bias = oa.normal((3 * Cc,), seed=74)
        want_qkv = _oracle_matmul(x, w, bias)
        assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        dx, dw, db = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf.from_numpy(bias)
        dq, do = pa.DevBuf(B * Cc * 4), pa.DevBuf(B * Cc * 4)
        pa.check(eng.qkv_append(0, dx.ptr, Cc, dw.ptr, db.ptr, dq.ptr, Cc), "qkv_append")
        pa.check(eng.decode(0, dq.ptr, Cc, do.ptr, Cc), "decode")
        eng.sync()
        q = dq.download((B, Cc))
        assert_close_gemm(q, want_qkv[:, :Cc], "q")
        k, v = eng.read_pool_rows(0, eng.slot_mapping())
        assert_close_gemm(k, want_qkv[:, Cc:2 * Cc], "k in the page slots")
        assert_close_gemm(v, want_qkv[:, 2 * Cc:], "v in the page slots")
        for s in range(B):
            orc.add_to_cache(want_qkv[s][None, None, :], 1, 1, 1, prompt=s)
        want = orc.decode_batch(sc.seq_ids, NH, want_qkv[:, :Cc])
        assert_close_gemm(do.download((B, Cc)), want, "decode after fused qkv append")
    finally:
        sc.close()


@pytest.mark.parametrize("split", [0, 1, 2, 3, 4, 8, 16, -2, -4])
@pytest.mark.parametrize("B,NH,hs", [(70, 12, 64), (130, 25, 64), (3, 2, 64), (20, 4, 64)])
def test_qkv_append_split_k(B, NH, hs, split):
    """Tensor-core GEMM with K split over several CTAs per tile -- split > 0: partial tiles through
    the L2 workspace, every CTA of a tile reducing its share of the columns in split order (0 = the
    automatic choice); split < 0: a cluster of 2/4 CTAs reducing through distributed shared memory in
    rank order.  Same result within the path tolerance, and bit-identical between two runs
    (deterministic reduction)."""
    bs = 16
    Cc = NH * hs
    sc = Scenario(NH, hs, bs, [3] * B, seed=95, extra_blocks=B + 8)
    try:
        eng = sc.eng
        x = oa.normal((B, Cc), seed=96)
        w = (oa.normal((3 * Cc, Cc), seed=97) * np.float32(1.0 / np.sqrt(Cc))).astype(np.float32)
        bias = oa.normal((3 * Cc,), seed=98)
        want = _oracle_matmul(x, w, bias)
        eng.tune(pa.PA_TUNE_GEMM_PATH, 2)
        eng.tune(pa.PA_TUNE_GEMM_SPLIT_K, split)
        dx, dw, db, dq = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf.from_numpy(bias), pa.DevBuf(B * Cc * 4)
        runs = []
        for _ in range(2):
            assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            pa.check(eng.qkv_append(0, dx.ptr, Cc, dw.ptr, db.ptr, dq.ptr, Cc), "qkv_append")
            eng.sync()
            k, v = eng.read_pool_rows(0, eng.slot_mapping())
            runs.append(np.concatenate([dq.download((B, Cc)), k, v], axis=1))
            pa.check(eng.step_rollback(), "rollback")
        assert_close_gemm(runs[0], want, f"split-K {split}")
        assert np.array_equal(runs[0].view(np.uint32), runs[1].view(np.uint32)), "split-K reduction is not deterministic"
    finally:
        sc.close()


@pytest.mark.parametrize("B,NH,hs", [(1, 12, 64), (4, 25, 64), (3, 3, 20), (2, 8, 128)])
def test_qkv_append_small_batch_gemv(B, NH, hs):
    """<= 4 new tokens: the weight-streaming GEMV kernel (auto-selected), with the KV scatter epilogue."""
    bs = 16
    Cc = NH * hs
    sc = Scenario(NH, hs, bs, [7] * B, seed=195, extra_blocks=B + 8)
    try:
        eng = sc.eng
        x = oa.normal((B, Cc), seed=196)
        w = (oa.normal((3 * Cc, Cc), seed=197) * np.float32(1.0 / np.sqrt(Cc))).astype(np.float32)
        bias = oa.normal((3 * Cc,), seed=198)
        want = _oracle_matmul(x, w, bias)
        for path in (0, 4):
            eng.tune(pa.PA_TUNE_GEMM_PATH, path)
            assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            dx, dw, db, dq = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf.from_numpy(bias), pa.DevBuf(B * Cc * 4)
            pa.check(eng.qkv_append(0, dx.ptr, Cc, dw.ptr, db.ptr, dq.ptr, Cc), "qkv_append")
            eng.sync()
            k, v = eng.read_pool_rows(0, eng.slot_mapping())
            assert_close_gemm(np.concatenate([dq.download((B, Cc)), k, v], axis=1), want, f"gemv path {path}")
            pa.check(eng.step_rollback(), "rollback")
    finally:
        sc.close()


def test_qkv_append_plain_tf32_has_its_own_tolerance():
    """PA_TUNE_GEMM_PATH=3: one TF32 MMA per k-step (reduced precision, opt-in): ~1e-3 of max|ref|."""
    NH, hs, bs, B = 12, 64, 16, 70
    Cc = NH * hs
    sc = Scenario(NH, hs, bs, [5] * B, seed=75, extra_blocks=B + 8)
    try:
        eng = sc.eng
        x = oa.normal((B, Cc), seed=76)
        w = (oa.normal((3 * Cc, Cc), seed=77) * np.float32(1.0 / np.sqrt(Cc))).astype(np.float32)
        want = _oracle_matmul(x, w, None)
        eng.tune(pa.PA_TUNE_GEMM_PATH, 3)
        assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        dx, dw, dq = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf(B * Cc * 4)
        pa.check(eng.qkv_append(0, dx.ptr, Cc, dw.ptr, None, dq.ptr, Cc), "qkv_append")
        eng.sync()
        err = np.abs(dq.download((B, Cc)).astype(np.float64) - want[:, :Cc]).max() / np.abs(want).max()
        assert 1e-6 < err <= 5e-3, err      # visibly TF32, and inside the TF32 tolerance
    finally:
        sc.close()


@pytest.mark.parametrize("path", [1, 2], ids=["simt", "tcgen05-3xtf32"])
def test_qkv_append_prefill_chunk(path):
    """Several new tokens per sequence (prompt chunk): every token's K/V lands in its own slot."""
    NH, hs, bs = 4, 64, 16
    Cc = NH * hs
    before, n_new = [0, 30, 7], [40, 3, 25]
    sc = Scenario(NH, hs, bs, before, seed=81, extra_blocks=16, max_batch_tokens=sum(n_new))
    try:
        eng = sc.eng
        eng.tune(pa.PA_TUNE_GEMM_PATH, path)
        ntok = sum(n_new)
        x = oa.normal((ntok, Cc), seed=82)
        w = (oa.normal((3 * Cc, Cc), seed=83) * np.float32(1.0 / np.sqrt(Cc))).astype(np.float32)
        want = _oracle_matmul(x, w, None)
        assert eng.step_begin(sc.seq_ids, n_new) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        dx, dw = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w)
        dq = pa.DevBuf(ntok * Cc * 4)
        pa.check(eng.qkv_append(0, dx.ptr, Cc, dw.ptr, None, dq.ptr, Cc), "qkv_append")
        eng.sync()
        assert_close_gemm(dq.download((ntok, Cc)), want[:, :Cc], "q")
        k, v = eng.read_pool_rows(0, eng.slot_mapping())
        assert_close_gemm(k, want[:, Cc:2 * Cc], "k")
        assert_close_gemm(v, want[:, 2 * Cc:], "v")
    finally:
        sc.close()


@pytest.mark.parametrize("device_buffers", [False, True], ids=["host-buffers", "device-buffers"])
@pytest.mark.parametrize("B,T,Cc", [(2, 5, 24), (3, 64, 768), (1, 1, 100)])
def test_compat_matmul_forward_and_cached(B, T, Cc, device_buffers):
    """The reference's own names and (B,T,OC) layout; matmul_cached leaves the K/V columns of the
    rows before the last one untouched (paged_infer.c:117-160)."""
    lib = pa.load()
    x = oa.normal((B * T, Cc), seed=91)
    w = (oa.normal((3 * Cc, Cc), seed=92) * np.float32(1.0 / np.sqrt(Cc))).astype(np.float32)
    bias = oa.normal((3 * Cc,), seed=93)
    for name, cached in (("matmul_forward", False), ("matmul_cached", True)):
        sentinel = np.float32(-7.25)
        out = np.full((B * T, 3 * Cc), sentinel, dtype=np.float32)
        if device_buffers:
            bufs = [pa.DevBuf.from_numpy(a) for a in (out, x, w, bias)]
            getattr(lib, name)(bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr, B, T, Cc, 3 * Cc)
            got = bufs[0].download(out.shape)
        else:
            getattr(lib, name)(out.ctypes.data, x.ctypes.data, w.ctypes.data, bias.ctypes.data, B, T, Cc, 3 * Cc)
            got = out
        want = np.full((B * T, 3 * Cc), sentinel, dtype=np.float32)
        ol = oa.load_oracle()
        fn = ol.orc_matmul_cached if cached else ol.orc_matmul_forward
        fn(oa.fptr(want), oa.fptr(x), oa.fptr(w), oa.fptr(bias), B, T, Cc, 3 * Cc)
        untouched = want == sentinel
        assert np.array_equal(got[untouched], want[untouched]), f"{name}: wrote outside its columns"
        assert_close_gemm(got, want, name)


@pytest.mark.parametrize("M,N,K", [(64, 768, 3072), (200, 768, 768), (7, 128, 2048), (64, 100, 1024), (130, 1600, 1600),
                                   (64, 3072, 768), (33, 64, 4096)])
def test_matmul_bias_auto_split_k(M, N, K):
    """pa_matmul_bias on device pointers at decode-step shapes: the tensor-core GEMM picks up to 16
    K-splits per tile by itself (workspace reduction).  Within the path tolerance of the oracle's
    matmul_forward, bit-identical across runs, and the workspace counters are back at rest (a third
    launch with another shape on the same stream still gives the right answer)."""
    lib = pa.load()
    x = oa.normal((M, K), seed=301)
    w = (oa.normal((N, K), seed=302) * np.float32(1.0 / np.sqrt(K))).astype(np.float32)
    bias = oa.normal((N,), seed=303)
    want = _oracle_matmul(x, w, bias)
    dx, dw, db, do = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf.from_numpy(bias), pa.DevBuf(M * N * 4)
    runs = []
    for _ in range(3):
        pa.check(lib.pa_memset(do.ptr, 0xff, M * N * 4, None), "memset")
        pa.check(lib.pa_matmul_bias(dx.ptr, K, dw.ptr, db.ptr, do.ptr, N, M, N, K, None), "pa_matmul_bias")
        pa.check(lib.pa_device_sync(), "sync")
        runs.append(do.download((M, N)))
    assert_close_gemm(runs[0], want, f"auto split-K {M}x{N}x{K}")
    assert np.array_equal(runs[0].view(np.uint32), runs[1].view(np.uint32))
    assert np.array_equal(runs[0].view(np.uint32), runs[2].view(np.uint32))
    # another shape on the same stream right after: the tile counters were reset
    x2 = oa.normal((M, 768), seed=304)
    w2 = (oa.normal((768, 768), seed=305) * np.float32(1.0 / np.sqrt(768))).astype(np.float32)
    want2 = _oracle_matmul(x2, w2, None)
    dx2, dw2, do2 = pa.DevBuf.from_numpy(x2), pa.DevBuf.from_numpy(w2), pa.DevBuf(M * 768 * 4)
    pa.check(lib.pa_matmul_bias(dx2.ptr, 768, dw2.ptr, None, do2.ptr, 768, M, 768, 768, None), "pa_matmul_bias")
    pa.check(lib.pa_device_sync(), "sync")
    assert_close_gemm(do2.download((M, 768)), want2, "second shape")


def test_split_k_with_two_streams_live():
    """Two handles (two streams) of one device both running workspace-split projections: from the second
    stream on the launches are cooperative (whole-grid residency guaranteed by the driver instead of
    assumed); each stream has its own workspace; results unchanged and bit-identical between the two."""
    NH, hs, bs, B = 12, 64, 16, 64
    Cc = NH * hs
    scs = [Scenario(NH, hs, bs, [3] * B, seed=500, extra_blocks=B + 8) for _ in range(2)]
    try:
        x = oa.normal((B, Cc), seed=501)
        w = (oa.normal((3 * Cc, Cc), seed=502) * np.float32(1.0 / np.sqrt(Cc))).astype(np.float32)
        bias = oa.normal((3 * Cc,), seed=503)
        want = _oracle_matmul(x, w, bias)
        outs = []
        bufs = [(pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf.from_numpy(bias), pa.DevBuf(B * Cc * 4)) for _ in scs]
        for rep in range(3):
            for sc, (dx, dw, db, dq) in zip(scs, bufs):
                eng = sc.eng
                eng.tune(pa.PA_TUNE_GEMM_PATH, 2)
                assert eng.step_begin(sc.seq_ids, [1] * B) == 0, pa.last_error()
                pa.check(eng.upload(), "upload")
                pa.check(eng.qkv_append(0, dx.ptr, Cc, dw.ptr, db.ptr, dq.ptr, Cc), "qkv_append")
            for sc, (dx, dw, db, dq) in zip(scs, bufs):
                eng = sc.eng
                eng.sync()
                k, v = eng.read_pool_rows(0, eng.slot_mapping())
                outs.append(np.concatenate([dq.download((B, Cc)), k, v], axis=1))
                pa.check(eng.step_rollback(), "rollback")
        for o in outs:
            assert_close_gemm(o, want, "two streams")
            assert np.array_equal(o.view(np.uint32), outs[0].view(np.uint32))
    finally:
        for sc in scs:
            sc.close()


@pytest.mark.timeout(180)
def test_split_k_spin_wait_beside_a_busy_foreign_stream():
    """the CTAs of a workspace-split projection wait for each other under a NON-cooperative launch, and
    the library only counts its own streams when it decides to launch cooperatively.  A caller's foreign stream that
    keeps the device busy (here: back-to-back device-side fills of a 1 GiB buffer on a stream the library knows
    nothing about -- 1184-CTA grids that hold SM slots but always finish) must delay the waiting CTAs, never deadlock
    them: a chain of 200 split projections queued beside it finishes, with bit-identical results to a quiet device.
    (pytest-timeout turns a hang into a failure.)"""
    lib = pa.load()
    M, N, K = 64, 768, 3072          # fcproj at 64 rows: 12 tiles x 12 splits = 144 spinning CTAs
    x = oa.normal((M, K), seed=801)
    w = (oa.normal((N, K), seed=802) * np.float32(1.0 / np.sqrt(K))).astype(np.float32)
    bias = oa.normal((N,), seed=803)
    want = _oracle_matmul(x, w, bias)
    dx, dw, db, do = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf.from_numpy(bias), pa.DevBuf(M * N * 4)
    ours, foreign = lib.pa_stream_create(), lib.pa_stream_create()
    junk_floats = 1 << 28
    junk = lib.pa_dev_alloc(junk_floats * 4)
    try:
        pa.check(lib.pa_matmul_bias(dx.ptr, K, dw.ptr, db.ptr, do.ptr, N, M, N, K, ours), "quiet")
        pa.check(lib.pa_stream_sync(ours), "sync")
        quiet = do.download((M, N))
        assert_close_gemm(quiet, want, "quiet device")
        for rep in range(40):
            pa.check(lib.pa_fill_normal(junk, junk_floats, 1.0, 0.0, rep, foreign), "foreign fill")
        for rep in range(200):
            pa.check(lib.pa_matmul_bias(dx.ptr, K, dw.ptr, db.ptr, do.ptr, N, M, N, K, ours), "busy")
        pa.check(lib.pa_stream_sync(ours), "sync ours")
        busy = do.download((M, N))
        pa.check(lib.pa_stream_sync(foreign), "sync foreign")
        assert np.array_equal(busy.view(np.uint32), quiet.view(np.uint32))
    finally:
        lib.pa_dev_free(junk)
        lib.pa_stream_destroy(ours)
        lib.pa_stream_destroy(foreign)
